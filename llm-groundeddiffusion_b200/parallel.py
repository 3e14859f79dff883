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


def _parse_cpu_list(s):
    out = []
    for part in s.split(","):
        a, _, b = part.partition("-")
        out.extend(range(int(a), int(b or a) + 1))
    return out


def gpu_cpu_affinity():
    """{gpu index: [cpu ids]} from `nvidia-smi topo -m` (the 'CPU Affinity' column); {} when unavailable"""
    import subprocess
    try:
        txt = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
    except Exception:
        return {}
    return parse_topo(txt)


def parse_topo(txt):
    """the GPU rows of `nvidia-smi topo -m`: first token that is a CPU list (digits, '-' and ',') after the link columns"""
    import re
    aff = {}
    for line in txt.splitlines():
        m = re.match(r"^GPU(\d+)\s", line)
        if not m:
            continue
        for tok in re.split(r"\s+", line.strip())[1:]:
            if re.fullmatch(r"\d+(-\d+)?(,\d+(-\d+)?)*", tok) and ("-" in tok or "," in tok):
                aff[int(m.group(1))] = _parse_cpu_list(tok)
                break
    return aff


def pin_host_threads(local_rank, ranks_on_node, max_threads=8):
    """One process per GPU: keep each rank's host work (latent composition, table building, launch loop) on its own
    slice of the cores that are local to its GPU, and cap torch's intra-op pool, so that N independent replicas do not
    contend for the same cores (round-1 scaling loss was host contention: eight ranks x all-core thread pools, ranks
    not NUMA-local).  Single-rank runs keep the full affinity (the CPU baseline uses all cores).
    Returns a short description for the bench line."""
    ncpu = os.cpu_count() or 1
    if ranks_on_node <= 1:
        torch.set_num_threads(min(16, ncpu))
        return {"pinned": False, "torch_threads": torch.get_num_threads()}
    aff = gpu_cpu_affinity()
    try:
        avail = sorted(os.sched_getaffinity(0))
    except AttributeError:
        avail = list(range(ncpu))
    mine = [c for c in aff.get(local_rank, avail) if c in set(avail)] or avail
    peers = [r for r in range(ranks_on_node) if aff.get(r, avail) == aff.get(local_rank, avail)] or [local_rank]
    k = peers.index(local_rank) if local_rank in peers else 0
    per = max(1, len(mine) // len(peers))
    cpus = mine[k * per:(k + 1) * per] or mine
    try:
        os.sched_setaffinity(0, cpus)
    except Exception:
        pass
    torch.set_num_threads(max(1, min(max_threads, len(cpus))))
    return {"pinned": True, "cpus": f"{cpus[0]}-{cpus[-1]} ({len(cpus)})", "torch_threads": torch.get_num_threads()}
