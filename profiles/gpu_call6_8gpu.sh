#!/bin/bash
# 8-GPU run of the bench (one rank per GPU, NCCL weight broadcast only): weak scaling at N = 8
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus 8 --steps 2 --warmup 2 --no-roofline --no-mode-a > gpurun_out/c6_bench_8gpu.json 2> gpurun_out/c6_bench_8gpu.err
echo "8gpu rc=$?"
tail -c 500 gpurun_out/c6_bench_8gpu.json; tail -4 gpurun_out/c6_bench_8gpu.err
