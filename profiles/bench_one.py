"""single-kernel driver for ncu --set full captures: python profiles/bench_one.py <case>"""
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
case = sys.argv[1] if len(sys.argv) > 1 else "proj320"
torch.manual_seed(0)
if case == "proj320":
    M, N, K = 65536, 320, 320
    x, w, res = torch.randn(M, K, device=dev).half(), torch.randn(N, K, device=dev).half(), torch.randn(M, N, device=dev).half()
    for _ in range(3):
        ops.linear(x, w, None, res)
elif case == "proj320_nores":
    M, N, K = 65536, 320, 320
    x, w = torch.randn(M, K, device=dev).half(), torch.randn(N, K, device=dev).half()
    for _ in range(3):
        ops.linear(x, w)
elif case == "toq":
    M, N, K = 2048, 1280, 1280
    x, w = torch.randn(M, K, device=dev).half(), torch.randn(N, K, device=dev).half()
    for _ in range(3):
        ops.linear(x, w)
elif case == "conv320":
    x = torch.randn(16, 64, 64, 320, device=dev).half()
    w = torch.randn(320, 9, 320, device=dev).half()
    for _ in range(3):
        ops.conv3x3(x, w)
elif case == "attn":
    B, heads, d, n = 16, 8, 40, 4096
    q, k, vt = ops.alloc_head_slabs(B, heads, d, n, n, dev)
    q.normal_(); k.normal_(); vt.normal_()
    q[:, :, d:] = 0
    k[:, :, d:] = 0
    for _ in range(3):
        ops.attention_fwd(q, k, vt, B, heads, n, n, d, d ** -0.5)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.attention_fwd(q, k, vt, B, heads, n, n, d, d ** -0.5)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("attn d=40 n=4096 B=16 ms", ms, "TF/s", 4.0 * n * n * d * heads * B / ms / 1e9)
elif case == "qkv320":
    B, heads, d, n, C = 16, 8, 40, 4096, 320
    x, w = torch.randn(B * n, C, device=dev).half(), torch.randn(3 * C, C, device=dev).half()
    q, k, vt = ops.alloc_head_slabs(B, heads, d, n, n, dev)
    for _ in range(3):
        ops.project_heads(x, w, n, heads, d, 0, q, k, vt)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.project_heads(x, w, n, heads, d, 0, q, k, vt)
    e1.record()
    torch.cuda.synchronize()
    print("qkv320 ms", e0.elapsed_time(e1) / 10)
elif case in ("toout320", "geglu320"):
    M, N, K = 65536, 320, 1280
    if case == "geglu320":
        N, K = 2560, 320
    x, w = torch.randn(M, K, device=dev).half(), torch.randn(N, K, device=dev).half()
    res = torch.randn(M, N, device=dev).half()
    if case == "geglu320":
        w_il, b_il = ops.geglu_interleave(w, torch.zeros(N, device=dev))
        fn = lambda: ops.linear_geglu(x, w_il, b_il)
    else:
        fn = lambda: ops.linear(x, w, None, res)
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(case, "ms", ms, "TF/s", 2.0 * M * N * K / ms / 1e9)
elif case == "attn_bwd":
    B, heads, d, n = 8, 8, 40, 4096
    dp, d16 = ops.round_dp(d), ops.round_d16(d)
    rm = lambda: torch.randn(B * heads, n, dp, device=dev).half()
    tr = lambda: torch.randn(B * heads, d16, n, device=dev).half()
    q, k, v, dO, qt, kt, dOt = rm(), rm(), rm(), rm(), tr(), tr(), tr()
    for t in (q, k, v, dO):
        t[:, :, d:] = 0
    vt = v[:, :, :d16].transpose(1, 2).contiguous()
    out, lse = ops.attention_fwd(q, k, vt, B, heads, n, n, d, d ** -0.5, want_lse=True)
    do_tok = torch.randn(B * n, heads * d, device=dev).half()
    for _ in range(2):
        ops.attention_bwd(q, k, v, dO, qt, kt, dOt, lse, out, do_tok, B, heads, n, n, d, d ** -0.5)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.attention_bwd(q, k, v, dO, qt, kt, dOt, lse, out, do_tok, B, heads, n, n, d, d ** -0.5)
    e1.record()
    torch.cuda.synchronize()
    print("attn_bwd d=40 n=4096 B=8 ms", e0.elapsed_time(e1) / 5)
torch.cuda.synchronize()
print("done", case)
