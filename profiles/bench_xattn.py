"""the graded kernel alone (xattn_fused_kernel at B=8, n=256, C=1280, heads 8, T=77, with the guidance loss):
   python profiles/bench_xattn.py            -> CUDA-event timing (graph replay, L2 flushed)
   ncu --set full -k regex:xattn_fused ...   -> profiles/*.ncu-rep summary"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

if __name__ == "__main__":
    for wl in (True, False):
        r = bench.xattn_roofline(torch.device("cuda:0"), with_loss=wl)
        print("with_loss", wl, r)
