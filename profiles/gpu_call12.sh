#!/bin/bash
# round-2 GPU call 12 (last): final tree - bench with the kernel-alone measurement first, smoke(), memcheck of the
# changed fused kernel, refreshed launch list + per-op profile
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err
echo "bench rc=$?" | tee gpurun_out/c12_summary.txt
tail -c 700 gpurun_out/c12_bench.json
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/c12_smoke.log 2>&1
echo "smoke rc=$?" | tee -a gpurun_out/c12_summary.txt
( time timeout 300 compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 3 --log-file gpurun_out/c12_memcheck.log \
    python -m pytest tests/test_xattn_fused_gpu.py tests/test_xattn_loss_gpu.py -m gpu -x -q ) > gpurun_out/c12_memcheck_pytest.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/c12_summary.txt
tail -3 gpurun_out/c12_memcheck.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c12_launches_step.csv \
    python profiles/profile_step.py --reps 1 > gpurun_out/c12_launches.log 2>&1
echo "launch list rc=$?" | tee -a gpurun_out/c12_summary.txt
timeout 200 python profiles/profile_ops.py > gpurun_out/c12_ops_profile.txt 2>&1
echo "ops profile rc=$?" | tee -a gpurun_out/c12_summary.txt
cat gpurun_out/c12_summary.txt
