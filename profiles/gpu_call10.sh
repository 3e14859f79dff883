#!/bin/bash
# round-2 GPU call 10: fixed-order GroupNorm reduction - repeatability of the gradient test, suite, bench
mkdir -p gpurun_out
: > gpurun_out/c10_repeat.txt
for i in 1 2 3 4; do
  python -m pytest tests/test_unet_gpu.py -q -s -k "guidance_gradient_matches or bit_reproducible" > gpurun_out/c10_repeat_$i.log 2>&1
  echo "run $i rc=$? $(grep 'grad rel-L2' gpurun_out/c10_repeat_$i.log | awk '{print $NF}' | tr '\n' ' ')" >> gpurun_out/c10_repeat.txt
done
cat gpurun_out/c10_repeat.txt
( time python -m pytest tests -m gpu -q -s ) > gpurun_out/c10_tests.log 2>&1
echo "tests rc=$?" | tee gpurun_out/c10_summary.txt
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err
echo "bench rc=$?" | tee -a gpurun_out/c10_summary.txt
tail -5 gpurun_out/c10_tests.log; cat gpurun_out/c10_summary.txt; tail -c 400 gpurun_out/c10_bench.json
