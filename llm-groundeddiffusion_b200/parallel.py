"""Multi-GPU plumbing: the path shards by image (independent prompt/layout pairs); the only
collective is the start-up broadcast of the frozen weights.  One process per GPU, torch.distributed (NCCL on B200,
gloo in the CPU tests)."""
import os

import torch
import torch.distributed as dist


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend="nccl", device=None):
    rank, world, _ = env_rank()
    if world > 1 and not dist.is_initialized():
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return rank, world


def broadcast_weights(shapes, make_on_rank0, device, src=0):
    """rank `src` materialises the state dict, every other rank allocates by shape and receives it"""
    rank = dist.get_rank() if dist.is_initialized() else 0
    if rank == src:
        w = make_on_rank0()
    else:
        w = {n: torch.empty(s, device=device) for n, s in shapes}
    if dist.is_initialized() and dist.get_world_size() > 1:
        for n in sorted(w):
            dist.broadcast(w[n], src=src)
    return w


def shard(items, rank, world):
    """static block partition of the prompt list (generate.py:23-25 does the same across processes)"""
    per = (len(items) + world - 1) // world
    return items[rank * per:(rank + 1) * per]


def max_over_ranks(value, device):
    t = torch.tensor([float(value)], device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
