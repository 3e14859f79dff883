"""world_size-2 gloo test of the N>1 plumbing (CPU): weight broadcast from rank 0, disjoint prompt shards, max-over-ranks
timing reduction."""
import os
import socket
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import lgd_b200  # noqa: F401
    from lgd_b200 import parallel, weights
    from lgd_b200.unet import UNetConfig
    r, w_ = parallel.init("gloo")
    cfg = UNetConfig(block_out_channels=(32, 64, 64, 64), cross_attention_dim=64)
    shapes = weights.parameter_shapes(cfg)
    w = parallel.broadcast_weights(shapes, lambda: weights.synthetic_weights(cfg, seed=7), "cpu")
    ck = sum(float(v.double().sum()) for v in w.values())
    mine = parallel.shard(list(range(16)), r, w_)
    mx = parallel.max_over_ranks(10.0 + r, "cpu")
    q.put((r, ck, mine, mx))
    torch.distributed.destroy_process_group()


def test_two_rank_broadcast_and_sharding():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(2))
    for p in ps:
        p.join(timeout=60)
    assert res[0][1] == res[1][1]                       # identical weights on both ranks
    assert res[0][2] == list(range(8)) and res[1][2] == list(range(8, 16))
    assert res[0][3] == res[1][3] == 11.0


def test_topo_parsing_and_rank_slices():
    """NUMA pinning input: the 'CPU Affinity' column of `nvidia-smi topo -m` (sample from the B200 box)"""
    import lgd_b200  # noqa: F401
    from lgd_b200 import parallel
    txt = "\n".join(["\tGPU0\tGPU1\tNIC0\tCPU Affinity\tNUMA Affinity\tGPU NUMA ID",
                     "GPU0\t X \tNV18\tSYS\t32-63,96-127\t1\t\tN/A",
                     "GPU1\tNV18\t X \tNODE\t0-31,64-95\t0\t\tN/A",
                     "NIC0\tSYS\tNODE\t X \t\t\t"])
    aff = parallel.parse_topo(txt)
    assert aff[0][:3] == [32, 33, 34] and len(aff[0]) == 64 and aff[1][-1] == 95
    assert parallel._parse_cpu_list("0-3,8") == [0, 1, 2, 3, 8]
