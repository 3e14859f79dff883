#!/bin/bash
# round-2 GPU call 5: suite with the persistent BoxDiff context, config-4 bench line, final default bench
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -s ) > gpurun_out/c5_tests.log 2>&1
echo "tests rc=$?" | tee gpurun_out/c5_summary.txt
python bench.py --workload boxdiff --steps 1 --warmup 2 > gpurun_out/c5_bench_config4.json 2> gpurun_out/c5_bench_config4.err
echo "bench config4 rc=$?" | tee -a gpurun_out/c5_summary.txt
B200_TIMING=1 python bench.py --steps 3 --warmup 3 > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err
echo "bench rc=$?" | tee -a gpurun_out/c5_summary.txt
ncu --set full --clock-control none -k regex:"gn_stats_kernel|gn_apply_kernel|ln_fwd_kernel|copy_rows_kernel|copy_cols_kernel" -c 16 \
    -o gpurun_out/c5_norm_kernels python profiles/profile_step.py --reps 1 > gpurun_out/c5_ncu_norm.log 2>&1
echo "ncu norm rc=$?" | tee -a gpurun_out/c5_summary.txt
tail -6 gpurun_out/c5_tests.log; cat gpurun_out/c5_summary.txt; tail -c 300 gpurun_out/c5_bench_config4.json; tail -4 gpurun_out/c5_bench_config4.err
