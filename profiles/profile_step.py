"""One CFG forward (batch 2B) + one guidance iteration (batch B) at SD1.4/1.5+GLIGEN shapes - the unit of work the bench
repeats - for ncu launch lists / full captures and CUDA-event phase timings.

    python profiles/profile_step.py [--batch 8] [--fuser 1] [--reps 3]
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python profiles/profile_step.py --reps 1
"""
import argparse
import os
import random
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def setup(batch=8, fuser=1, tiny=0):
    import lgd_b200  # noqa: F401
    from lgd_b200 import guidance as G, pipelines as P, weights as Wt
    from lgd_b200.unet import B200UNet, UNetConfig
    dev = torch.device("cuda:0")
    cfg = UNetConfig.tiny(gligen=True) if tiny else UNetConfig.sd15(gligen=True)
    net = B200UNet(cfg, Wt.synthetic_weights(cfg, 0, dev), dev)
    B, side = batch, 64
    g = torch.Generator().manual_seed(0)
    z = torch.randn(B, 4, side, side, generator=g).to(dev)
    text = torch.randn(2 * B, 77, 768, generator=g)
    kv = net.set_text(text)
    heads = 8
    kv_cond = lambda p: tuple(s[B * heads:] for s in kv.slabs[p])
    rng = random.Random(0)
    lay = []
    for b in range(B):
        bx, pos = [], []
        for o in range(4):
            w_, h_ = rng.uniform(0.2, 0.5), rng.uniform(0.2, 0.5)
            x0, y0 = rng.uniform(0, 1 - w_), rng.uniform(0, 1 - h_)
            bx.append([(x0, y0, x0 + w_, y0 + h_)])
            pos.append([2 * o + 1, 2 * o + 2])
        lay.append(G.SampleLayout(bx, pos, [p[-1] for p in pos]))
    spec = P.GuidanceSpec(layouts=lay, loss_scale=5, loss_threshold=0.0, max_iter=1, max_index_step=30, fg_weight=1.0,
                          bg_weight=4.0)
    losses = P.build_losses(net, spec, 0, side, side, dev)[0]
    gl = dict(boxes=torch.rand(2 * B, 30, 4, generator=g), masks=(torch.rand(2 * B, 30, generator=g) > 0.8).float(),
              positive_embeddings=torch.randn(2 * B, 30, 768, generator=g))
    objs = net.position_net(gl["boxes"], gl["masks"], gl["positive_embeddings"])
    t2 = torch.full((2 * B,), 500.0, device=dev)
    t1 = torch.full((B,), 500.0, device=dev)

    return dict(net=net, z=z, t2=t2, t1=t1, kv=kv, kv_cond=kv_cond, losses=losses, objs=objs, B=B, fuser=bool(fuser))


def run_phase(c, phase):
    net, B = c["net"], c["B"]
    if phase == "forward":
        net.forward(c["z"], c["t2"], c["kv"], rep=2, objs=c["objs"], fuser_on=c["fuser"])
    else:
        net.guidance_gradient(c["z"], c["t1"], c["kv_cond"], c["losses"], objs=c["objs"][:B * 30], fuser_on=c["fuser"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--fuser", type=int, default=1)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--tiny", type=int, default=0)
    a = ap.parse_args()
    c = setup(a.batch, a.fuser, a.tiny)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    for rep in range(a.reps):
        e = [ev() for _ in range(3)]
        e[0].record()
        run_phase(c, "forward")
        e[1].record()
        run_phase(c, "guidance")
        e[2].record()
        torch.cuda.synchronize()
        print(f"rep {rep}: CFG forward (batch {2 * c['B']}) {e[0].elapsed_time(e[1]):.2f} ms, guidance fwd+bwd (batch {c['B']}) "
              f"{e[1].elapsed_time(e[2]):.2f} ms")


if __name__ == "__main__":
    main()
