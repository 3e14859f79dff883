"""Per-op CUDA-event timing of one CFG forward (batch 2B) + one guidance iteration (batch B) at SD1.5+GLIGEN shapes:
every C-ABI call the step makes is bracketed by events (eager launches, no graph), grouped by (entry point, shape).
GEMM rows carry M, N, K and the achieved TFLOP/s.  Used to pick the next kernel to work on.

    python profiles/profile_ops.py [--batch 8] > profiles/rN_ops.txt
"""
import argparse
import collections
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "profiles"))


class Recorder:
    def __init__(self):
        self.rows = []
        self.on = False

    def wrap(self, fn, keyfn):
        def inner(*a, **k):
            if not self.on:
                return fn(*a, **k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            self.rows.append((keyfn(*a, **k), e0, e1))
            return r
        return inner


class LibProxy:
    def __init__(self, lib, rec):
        self._lib, self._rec, self._cache = lib, rec, {}

    def __getattr__(self, name):
        if name not in self._cache:
            f = getattr(self._lib, name)
            def keyfn(*a, _n=name, **k):
                if "attention" in _n or "xattn" in _n:     # shape ints (B, heads, nq, nk, ..., d) tell the calls apart
                    import ctypes
                    return (_n,) + tuple(x.value for x in a if isinstance(x, ctypes.c_int))
                return (_n,)
            self._cache[name] = self._rec.wrap(f, keyfn) if name.startswith("b200lmd_") and name not in ("b200lmd_gemm", "b200lmd_round_dp", "b200lmd_round_d16") else f
        return self._cache[name]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    a = ap.parse_args()
    import lgd_b200  # noqa: F401
    from lgd_b200 import unet as U
    import profile_step as PS
    rec = Recorder()
    raw_gemm = U.gemm

    def gemm_key(A, a_geom, W, N, wtaps, grid, taps, **k):
        M = grid[0] * grid[1] * grid[2]
        K = a_geom[4] * len(taps)
        return ("gemm", M, N, K, len(taps), k.get("mode", 0), "res" if k.get("residual") is not None else "",
                "f32" if k.get("out_f32") is not None else "", "acc" if k.get("accumulate") else "")

    U.gemm = rec.wrap(raw_gemm, gemm_key)
    proxy = LibProxy(U.lib(), rec)
    U.lib = lambda: proxy
    from lgd_b200 import ops as O
    O.lib = lambda: proxy
    ctx = PS.setup(a.batch)
    for rep in range(3):
        rec.rows.clear()
        rec.on = rep == 2
        for phase in ("forward", "guidance"):
            start = len(rec.rows)
            PS.run_phase(ctx, phase)
            torch.cuda.synchronize()
            if rec.on:
                agg = collections.OrderedDict()
                for key, e0, e1 in rec.rows[start:]:
                    c = agg.setdefault(key, [0, 0.0])
                    c[0] += 1
                    c[1] += e0.elapsed_time(e1)
                tot = sum(v[1] for v in agg.values())
                print(f"== {phase}: {tot:.2f} ms in ops (event-bracketed, eager)")
                for key, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                    extra = ""
                    if key[0] == "gemm":
                        fl = 2.0 * key[1] * key[2] * key[3] * cnt
                        extra = f"  {fl / ms / 1e9:7.1f} TF/s"
                    print(f"{ms:8.3f} ms {100 * ms / tot:5.1f}%  x{cnt:<3d} {key}{extra}")


if __name__ == "__main__":
    main()
