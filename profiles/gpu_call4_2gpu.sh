#!/bin/bash
# 2-GPU sanity of the multi-rank path (NCCL weight broadcast, NUMA pinning, max-over-ranks timing)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/c4_topo_2gpu.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 1 --warmup 2 --no-roofline > gpurun_out/c4_bench_2gpu.json 2> gpurun_out/c4_bench_2gpu.err
echo "2gpu rc=$?"
tail -c 600 gpurun_out/c4_bench_2gpu.json; tail -5 gpurun_out/c4_bench_2gpu.err
