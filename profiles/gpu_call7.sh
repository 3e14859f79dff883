#!/bin/bash
# round-2 GPU call 7: suite incl. the full-width LMD+ golden test and the multi-row LayerNorm, per-op profile, bench
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -s ) > gpurun_out/c7_tests.log 2>&1
echo "tests rc=$?" | tee gpurun_out/c7_summary.txt
python profiles/profile_ops.py > gpurun_out/c7_ops_profile.txt 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err
echo "bench rc=$?" | tee -a gpurun_out/c7_summary.txt
tail -6 gpurun_out/c7_tests.log; cat gpurun_out/c7_summary.txt; grep layernorm gpurun_out/c7_ops_profile.txt | head -4; tail -c 300 gpurun_out/c7_bench.json
