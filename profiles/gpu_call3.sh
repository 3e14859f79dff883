#!/bin/bash
# round-2 GPU call 3: suite after the fused-kernel loss staging / GEGLU revert, timeline A/B, bench (VAE on), configs 3/4
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -s ) > gpurun_out/c3_tests.log 2>&1
echo "tests rc=$?" | tee gpurun_out/c3_summary.txt
python profiles/timeline_xattn.py > gpurun_out/c3_timeline.txt 2>&1
B200_TIMING=1 python bench.py --steps 2 --warmup 3 > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err
echo "bench rc=$?" | tee -a gpurun_out/c3_summary.txt
python bench.py --workload backward_guidance_sd21 --steps 1 --warmup 2 > gpurun_out/c3_bench_config3.json 2> gpurun_out/c3_bench_config3.err
echo "bench config3 rc=$?" | tee -a gpurun_out/c3_summary.txt
python bench.py --workload boxdiff --steps 1 --warmup 2 > gpurun_out/c3_bench_config4.json 2> gpurun_out/c3_bench_config4.err
echo "bench config4 rc=$?" | tee -a gpurun_out/c3_summary.txt
python profiles/profile_ops.py > gpurun_out/c3_ops_profile.txt 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c3_bench_reference.json 2> gpurun_out/c3_bench_reference.err
echo "bench reference rc=$?" | tee -a gpurun_out/c3_summary.txt
tail -6 gpurun_out/c3_tests.log; cat gpurun_out/c3_summary.txt; cat gpurun_out/c3_timeline.txt | grep "with_loss"; tail -c 400 gpurun_out/c3_bench.json
