#!/bin/bash
# round-2 GPU call 14: BASELINE config 3 / 4 bench lines on the final tree
mkdir -p gpurun_out
timeout 150 python bench.py --workload boxdiff --steps 2 --warmup 3 > gpurun_out/c14_bench_boxdiff.json 2> gpurun_out/c14_bench_boxdiff.err
echo "boxdiff rc=$?"
timeout 170 python bench.py --workload backward_guidance_sd21 --steps 2 --warmup 3 > gpurun_out/c14_bench_bg_sd21.json 2> gpurun_out/c14_bench_bg_sd21.err
echo "bg sd21 rc=$?"
head -c 250 gpurun_out/c14_bench_boxdiff.json; echo; head -c 250 gpurun_out/c14_bench_bg_sd21.json
