#!/bin/bash
# round-2 GPU call 4: suite with the device-side loop predicate, bench, config 3 / 4 bench lines
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -s ) > gpurun_out/c4_tests.log 2>&1
echo "tests rc=$?" | tee gpurun_out/c4_summary.txt
B200_TIMING=1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err
echo "bench rc=$?" | tee -a gpurun_out/c4_summary.txt
python bench.py --workload backward_guidance_sd21 --steps 1 --warmup 2 > gpurun_out/c4_bench_config3.json 2> gpurun_out/c4_bench_config3.err
echo "bench config3 rc=$?" | tee -a gpurun_out/c4_summary.txt
python bench.py --workload boxdiff --steps 1 --warmup 2 > gpurun_out/c4_bench_config4.json 2> gpurun_out/c4_bench_config4.err
echo "bench config4 rc=$?" | tee -a gpurun_out/c4_summary.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c4_smoke.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/c4_summary.txt
tail -6 gpurun_out/c4_tests.log; cat gpurun_out/c4_summary.txt; tail -c 300 gpurun_out/c4_bench.json; tail -c 300 gpurun_out/c4_bench_config3.json; tail -c 300 gpurun_out/c4_bench_config4.json
