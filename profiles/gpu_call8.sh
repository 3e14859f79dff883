#!/bin/bash
# round-2 GPU call 8: suite + per-op profile + bench after templating gemm2 on its epilogue mode
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -s ) > gpurun_out/c8_tests.log 2>&1
echo "tests rc=$?" | tee gpurun_out/c8_summary.txt
python profiles/profile_ops.py > gpurun_out/c8_ops_profile.txt 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/c8_bench.json 2> gpurun_out/c8_bench.err
echo "bench rc=$?" | tee -a gpurun_out/c8_summary.txt
tail -6 gpurun_out/c8_tests.log; cat gpurun_out/c8_summary.txt; head -12 gpurun_out/c8_ops_profile.txt | cut -c1-110; tail -c 300 gpurun_out/c8_bench.json
