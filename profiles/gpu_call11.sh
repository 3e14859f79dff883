#!/bin/bash
# round-2 GPU call 11: fused kernel with the loss path shortened (early residual/bias, prefetch under the cluster
# barrier, vector loads, deferred ticket): suite, timeline, bench (launch floor / back-to-back), fresh ncu capture
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_xattn_fused_gpu.py tests/test_xattn_loss_gpu.py -q -x > gpurun_out/c11_fused_tests.log 2>&1
echo "fused tests rc=$?" | tee gpurun_out/c11_summary.txt
tail -3 gpurun_out/c11_fused_tests.log
python profiles/timeline_xattn.py > gpurun_out/c11_timeline.txt 2>&1
echo "timeline rc=$?" | tee -a gpurun_out/c11_summary.txt
grep with_loss gpurun_out/c11_timeline.txt
( time python -m pytest tests -m gpu -q -s ) > gpurun_out/c11_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/c11_summary.txt
tail -4 gpurun_out/c11_tests.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c11_bench.json 2> gpurun_out/c11_bench.err
echo "bench rc=$?" | tee -a gpurun_out/c11_summary.txt
tail -c 900 gpurun_out/c11_bench.json
timeout 240 ncu --set full --clock-control none --import-source on -k regex:xattn_fused_kernel -s 3 -c 1 -f -o gpurun_out/c11_xattn_fused \
    python profiles/bench_xattn.py > gpurun_out/c11_ncu_xattn.log 2>&1
echo "ncu rc=$?" | tee -a gpurun_out/c11_summary.txt
ls -la gpurun_out/c11_xattn_fused.ncu-rep
