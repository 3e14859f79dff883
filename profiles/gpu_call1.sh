#!/bin/bash
# round-2 GPU call 1: new parity tests, whole suite, bench with phase timing, compute-sanitizer memcheck
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_gpu.txt
nproc >> gpurun_out/c1_gpu.txt
nvidia-smi topo -m >> gpurun_out/c1_gpu.txt 2>&1
( time python -m pytest tests/test_fullwidth_gpu.py tests/test_layout_gpu.py tests/test_adapter_gpu.py tests/test_boxdiff_gpu.py tests/test_plugins_gpu.py -m gpu -q -s ) > gpurun_out/c1_newtests.log 2>&1
echo "new tests rc=$?" | tee -a gpurun_out/c1_summary.txt
( time python -m pytest tests -m gpu -q --deselect tests/test_fullwidth_gpu.py --deselect tests/test_layout_gpu.py --deselect tests/test_adapter_gpu.py --deselect tests/test_boxdiff_gpu.py --deselect tests/test_plugins_gpu.py ) > gpurun_out/c1_oldtests.log 2>&1
echo "old tests rc=$?" | tee -a gpurun_out/c1_summary.txt
B200_TIMING=1 python bench.py --steps 1 --warmup 2 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
echo "bench rc=$?" | tee -a gpurun_out/c1_summary.txt
( time timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 --log-file gpurun_out/c1_memcheck.log \
    python -m pytest tests/test_unet_gpu.py tests/test_pipeline_gpu.py tests/test_xattn_fused_gpu.py tests/test_xattn_loss_gpu.py -m gpu -x -q ) > gpurun_out/c1_memcheck_pytest.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/c1_summary.txt
tail -5 gpurun_out/c1_newtests.log; tail -3 gpurun_out/c1_oldtests.log; tail -3 gpurun_out/c1_memcheck_pytest.log; tail -c 1500 gpurun_out/c1_bench.json
