"""phase timeline of xattn_fused_kernel from in-kernel %globaltimer stamps (B=8, n=256, C=1280): per-CTA phase times;
slots 8..10 are the loss reduction's own stamps"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from lgd_b200._lib import lib  # noqa: E402

dev = torch.device("cuda:0")
dbg = torch.zeros(128 * 16, dtype=torch.int64, device=dev)
for wl, stage in ((False, 1), (True, 0), (True, 1)):
    lib().b200lmd_set_option(b"fused_loss_stage", ctypes.c_int(stage))
    dbg.zero_()
    lib().b200lmd_set_debug_buffer(ctypes.c_void_p(dbg.data_ptr()))
    r = bench.xattn_roofline(dev, with_loss=wl)
    torch.cuda.synchronize()
    t = dbg.view(128, 16).cpu().double()
    t0 = t[:, 0].min()
    names = ["start", "Q acc done", "S done", "O done", "pre-sync2", "post-sync2", "out acc done", "end",
             "loss run", "loss probs", "loss end"]
    print("with_loss", wl, "loss inputs staged in smem", bool(stage), "ms_per_op", r["ms_per_op"], "frac", r["frac"])
    for i, nme in enumerate(names):
        col = t[:, i]
        col = col[col > 0]
        if col.numel() == 0:
            continue
        rel = (col - t0) / 1e3
        print(f"  {nme:14s} n={col.numel():3d} min {rel.min():8.2f}  median {rel.median():8.2f}  max {rel.max():8.2f} us")
lib().b200lmd_set_debug_buffer(None)
