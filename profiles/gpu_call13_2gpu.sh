#!/bin/bash
# round-2 GPU call 13: the driver's N=2 command on the final tree (roofline micro-benchmark now runs first on rank 0)
mkdir -p gpurun_out
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus 2 --steps 1 --warmup 3 > gpurun_out/c13_bench_2gpu.json 2> gpurun_out/c13_bench_2gpu.err
echo "2gpu rc=$?"
tail -c 500 gpurun_out/c13_bench_2gpu.json; tail -4 gpurun_out/c13_bench_2gpu.err
