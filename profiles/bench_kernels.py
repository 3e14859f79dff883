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
    ms = timeit(lambda: ops.linear(x, w))
    print(f"{name}: M={M} N={N} K={K}: {ms * 1e3:.1f} us  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s")


def conv(B, H, C_in, C_out):
    x = torch.randn(B, H, H, C_in, device=dev).half()
    w = torch.randn(C_out, 9, C_in, device=dev).half()
    ms = timeit(lambda: ops.conv3x3(x, w))
    print(f"conv3x3 B={B} {H}x{H} {C_in}->{C_out}: {ms * 1e3:.1f} us  {2.0 * B * H * H * C_out * 9 * C_in / ms / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    attn(16, 8, 40, 4096)
    attn(16, 8, 80, 1024)
    attn(8, 8, 64, 4096)
    conv(16, 64, 320, 320)
    conv(16, 32, 640, 640)
    conv(16, 16, 1280, 1280)
    conv(16, 8, 1280, 1280)
    gemm(65536, 320, 320, "proj 320")
    gemm(65536, 2560, 320, "geglu-in 320")
    gemm(65536, 320, 1280, "ff-out 320")
    gemm(4096, 1280, 1280, "proj 1280")
    gemm(2048, 1280, 1280, "to_q guidance")
