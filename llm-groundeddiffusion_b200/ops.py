"""Thin tensor-level wrappers over the stand-alone C-ABI ops (used by tests and by the host-side mirrors)."""
import ctypes

import torch

from ._lib import check, cur_stream, lib, ptr

_f = ctypes.c_float
_i = ctypes.c_int


def linear(x, w, bias=None, residual=None, alpha=1.0, out=None, out_f32=False, accumulate=False):
    """x [M,K] fp16, w [N,K] fp16 -> [M,N] fp16 (models/attention_processor.py:127-142 etc.)"""
    M, K = x.shape
    N = w.shape[0]
    assert x.dtype == torch.float16 and w.dtype == torch.float16 and x.stride(1) == 1
    y = out if out is not None else torch.empty(M, N, device=x.device, dtype=torch.float16)
    y32 = torch.empty(M, N, device=x.device, dtype=torch.float32) if out_f32 else None
    check(lib().b200lmd_linear_f16(ptr(x), _i(x.stride(0)), ptr(w), ptr(bias), ptr(residual),
                                   _i(residual.stride(0) if residual is not None else 0), ptr(y), _i(y.stride(0)),
                                   ptr(y32), _i(M), _i(N), _i(K), _f(alpha), _i(int(accumulate)), cur_stream()))
    return (y, y32) if out_f32 else y


def geglu_interleave(w, bias):
    F2, K = w.shape
    F = F2 // 2
    w_il = torch.empty_like(w)
    b_il = torch.empty_like(bias)
    check(lib().b200lmd_geglu_interleave_w(ptr(w), ptr(w_il), _i(F), _i(K), cur_stream()))
    check(lib().b200lmd_geglu_interleave_b(ptr(bias), ptr(b_il), _i(F), cur_stream()))
    return w_il, b_il


def linear_geglu(x, w_il, bias_il, want_pre=False):
    M, K = x.shape
    F = w_il.shape[0] // 2
    y = torch.empty(M, F, device=x.device, dtype=torch.float16)
    pre = torch.empty(M, 2 * F, device=x.device, dtype=torch.float16) if want_pre else None
    check(lib().b200lmd_linear_geglu_f16(ptr(x), _i(x.stride(0)), ptr(w_il), ptr(bias_il), ptr(y), ptr(pre), _i(M),
                                         _i(F), _i(K), cur_stream()))
    return (y, pre) if want_pre else y


def conv3x3(x, w, bias=None, chan_add=None, residual=None, out_f32=False):
    """x [B,H,W,Cin] fp16 NHWC, w [Cout,9,Cin] fp16 -> [B,H,W,Cout]"""
    B, H, W, Cin = x.shape
    Cout = w.shape[0]
    y = torch.empty(B, H, W, Cout, device=x.device, dtype=torch.float16)
    y32 = torch.empty(B, H, W, Cout, device=x.device, dtype=torch.float32) if out_f32 else None
    check(lib().b200lmd_conv3x3_f16(ptr(x), ptr(w), ptr(bias), ptr(chan_add), ptr(residual), ptr(y), ptr(y32), _i(B),
                                    _i(H), _i(W), _i(Cin), _i(Cout), cur_stream()))
    return (y, y32) if out_f32 else y
