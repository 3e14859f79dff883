"""CUDA-event micro-benchmarks of individual kernels at the bench workload's shapes (device time, graph replay, L2
flushed between repetitions).  python profiles/bench_kernels.py"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lgd_b200  # noqa: E402
from lgd_b200 import ops  # noqa: E402
from lgd_b200._lib import lib  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


def attn(B, heads, d, n):
    q, k, vt = ops.alloc_head_slabs(B, heads, d, n, n, dev)
    q.normal_()
    k.normal_()
    vt.normal_()
    q[:, :, d:] = 0
    k[:, :, d:] = 0
    flops = 4.0 * n * n * d * B * heads
    for v2 in (0, 1):
        lib().b200lmd_set_option(b"attn_v2", ctypes.c_int(v2))
        ms = timeit(lambda: ops.attention_fwd(q, k, vt, B, heads, n, n, d, d ** -0.5))
        print(f"attention fwd B={B} heads={heads} d={d} n={n} v2={v2}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s (useful)")


def gemm(M, N, K, name):
    x = torch.randn(M, K, device=dev).half()
    w = torch.randn(N, K, device=dev).half()
    res = torch.randn(M, N, device=dev).half()
    for v2 in (1, 3):
        lib().b200lmd_set_option(b"gemm_v3", ctypes.c_int(int(v2 == 3)))
        ms = timeit(lambda: ops.linear(x, w, None, res))
        byts = 2.0 * (M * K + 2 * M * N + N * K)
        print(f"{name} v2={v2}: M={M} N={N} K={K}: {ms * 1e3:.1f} us  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s  "
              f"{byts / ms / 1e6:.0f} GB/s")


def geglu(M, F, K):
    x = torch.randn(M, K, device=dev).half()
    w = (torch.randn(2 * F, K, device=dev)).half()
    b = torch.randn(2 * F, device=dev)
    w_il, b_il = ops.geglu_interleave(w, b)
    for v2 in (0, 1):
        lib().b200lmd_set_option(b"gemm_v2", ctypes.c_int(v2))
        ms = timeit(lambda: ops.linear_geglu(x, w_il, b_il))
        print(f"geglu v2={v2}: M={M} F={F} K={K}: {ms * 1e3:.1f} us  {4.0 * M * F * K / ms / 1e9:.1f} TFLOP/s  "
              f"{2.0 * (M * K + M * F) / ms / 1e6:.0f} GB/s")


def conv(B, H, C_in, C_out):
    x = torch.randn(B, H, H, C_in, device=dev).half()
    w = torch.randn(C_out, 9, C_in, device=dev).half()
    for v2 in (1, 3):
        lib().b200lmd_set_option(b"gemm_v3", ctypes.c_int(int(v2 == 3)))
        ms = timeit(lambda: ops.conv3x3(x, w))
        print(f"conv3x3 v2={v2} B={B} {H}x{H} {C_in}->{C_out}: {ms * 1e3:.1f} us  "
              f"{2.0 * B * H * H * C_out * 9 * C_in / ms / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    attn(16, 8, 40, 4096)
    attn(16, 8, 80, 1024)
    conv(16, 64, 320, 320)
    conv(16, 32, 640, 640)
    conv(16, 16, 1280, 1280)
    conv(16, 8, 1280, 1280)
    gemm(65536, 320, 320, "proj 320")
    geglu(65536, 1280, 320)
    geglu(16384, 2560, 640)
    gemm(65536, 320, 1280, "ff-out 320")
    gemm(4096, 1280, 1280, "proj 1280")
    gemm(2048, 1280, 1280, "to_q guidance")
