#!/bin/bash
# round-2 GPU call 2: whole suite, bench (graph cache) with phase timing, fused-kernel timeline + ncu, launch list,
# per-op profile, memcheck / racecheck
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -s ) > gpurun_out/c2_tests.log 2>&1
echo "tests rc=$?" | tee gpurun_out/c2_summary.txt
B200_TIMING=1 python bench.py --steps 2 --warmup 3 > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
echo "bench rc=$?" | tee -a gpurun_out/c2_summary.txt
python bench.py --steps 1 --warmup 2 --vae 1 --no-roofline --no-cpu-baseline --no-mode-a > gpurun_out/c2_bench_vae.json 2> gpurun_out/c2_bench_vae.err
echo "bench vae rc=$?" | tee -a gpurun_out/c2_summary.txt
python profiles/timeline_xattn.py > gpurun_out/c2_timeline.txt 2>&1
python profiles/profile_ops.py > gpurun_out/c2_ops_profile.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:xattn_fused -s 3 -c 1 -o gpurun_out/c2_xattn_fused \
    python profiles/bench_xattn.py > gpurun_out/c2_ncu_xattn.log 2>&1
echo "ncu rc=$?" | tee -a gpurun_out/c2_summary.txt
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c2_launches_step.csv \
    python profiles/profile_step.py --reps 1 > gpurun_out/c2_launches.log 2>&1
( time timeout 600 compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 3 --log-file gpurun_out/c2_memcheck.log \
    python -m pytest tests/test_unet_gpu.py tests/test_pipeline_gpu.py tests/test_xattn_fused_gpu.py tests/test_boxdiff_gpu.py tests/test_plugins_gpu.py -m gpu -x -q ) > gpurun_out/c2_memcheck_pytest.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/c2_summary.txt
( time timeout 420 compute-sanitizer --tool racecheck --report-api-errors no --error-exitcode 3 --log-file gpurun_out/c2_racecheck.log \
    python -m pytest tests/test_xattn_fused_gpu.py -m gpu -x -q ) > gpurun_out/c2_racecheck_pytest.log 2>&1
echo "racecheck rc=$?" | tee -a gpurun_out/c2_summary.txt
tail -8 gpurun_out/c2_tests.log; cat gpurun_out/c2_summary.txt; tail -c 600 gpurun_out/c2_bench.json
