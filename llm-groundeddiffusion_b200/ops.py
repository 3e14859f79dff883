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


def round_dp(d):
    return lib().b200lmd_round_dp(_i(d))


def round_d16(d):
    return lib().b200lmd_round_d16(_i(d))


def alloc_head_slabs(B, heads, d, n_alloc_q, n_alloc_k, device):
    dp, d16 = round_dp(d), round_d16(d)
    q = torch.zeros(B * heads, n_alloc_q, dp, device=device, dtype=torch.float16)
    k = torch.zeros(B * heads, n_alloc_k, dp, device=device, dtype=torch.float16)
    vt = torch.zeros(B * heads, d16, n_alloc_k, device=device, dtype=torch.float16)
    return q, k, vt


def project_heads(x, w, rows_per_img, heads, d, which0, q=None, k=None, vt=None):
    """x [M,K] @ w[nproj*C, K]^T scattered into slabs (see include/b200lmd.h)"""
    M, K = x.shape
    N = w.shape[0]
    check(lib().b200lmd_project_heads_f16(ptr(x), _i(x.stride(0)), ptr(w), _i(M), _i(N), _i(K), _i(rows_per_img),
                                          _i(heads), _i(d), _i(which0), ptr(q), _i(q.shape[1] if q is not None else 0),
                                          ptr(k), _i(k.shape[1] if k is not None else 0), ptr(vt),
                                          _i(vt.shape[2] if vt is not None else 0), cur_stream()))


def attention_fwd(q, k, vt, B, heads, nq, nk, d, scale, want_lse=False):
    out = torch.empty(B * nq, heads * d, device=q.device, dtype=torch.float16)
    lse = torch.zeros(B * heads, q.shape[1], device=q.device, dtype=torch.float32) if want_lse else None
    check(lib().b200lmd_attention_fwd_f16(ptr(q), ptr(k), ptr(vt), ptr(out), _i(heads * d), ptr(lse), _i(B), _i(heads),
                                          _i(nq), _i(nk), _i(q.shape[1]), _i(k.shape[1]), _i(d), _f(scale),
                                          cur_stream()))
    return (out, lse) if want_lse else out


def groupnorm(x, gamma, beta, groups, eps, silu, want_sums=False):
    """x [B, n, C] fp16"""
    B, n, C = x.shape
    y = torch.empty_like(x)
    sums = torch.empty(B, groups, 2, device=x.device, dtype=torch.float32)
    check(lib().b200lmd_groupnorm_f16(ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(sums), _i(B), _i(n), _i(C),
                                      _i(groups), _f(eps), _i(int(silu)), cur_stream()))
    return (y, sums) if want_sums else y


def groupnorm_bwd(dy, x, sums, gamma, beta, groups, eps, silu, dx=None):
    B, n, C = x.shape
    acc = dx is not None
    if dx is None:
        dx = torch.empty_like(x)
    bsums = torch.empty(B, groups, 2, device=x.device, dtype=torch.float32)
    check(lib().b200lmd_groupnorm_bwd_f16(ptr(dy), ptr(x), ptr(sums), ptr(gamma), ptr(beta), ptr(dx), ptr(bsums), _i(B),
                                          _i(n), _i(C), _i(groups), _f(eps), _i(int(silu)), _i(int(acc)),
                                          cur_stream()))
    return dx


def layernorm(x, gamma, beta, eps=1e-5, want_stats=False):
    rows, C = x.shape
    y = torch.empty_like(x)
    stats = torch.empty(rows, 2, device=x.device, dtype=torch.float32) if want_stats else None
    check(lib().b200lmd_layernorm_f16(ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(stats), ctypes.c_longlong(rows),
                                      _i(C), _f(eps), cur_stream()))
    return (y, stats) if want_stats else y


def layernorm_bwd(dy, x, stats, gamma, dx=None):
    rows, C = x.shape
    acc = dx is not None
    if dx is None:
        dx = torch.empty_like(x)
    check(lib().b200lmd_layernorm_bwd_f16(ptr(dy), ptr(x), ptr(stats), ptr(gamma), ptr(dx), ctypes.c_longlong(rows),
                                          _i(C), _i(int(acc)), cur_stream()))
    return dx


def geglu_bwd(pre, dy):
    rows, F = dy.shape
    dpre = torch.empty_like(pre)
    check(lib().b200lmd_geglu_bwd_f16(ptr(pre), ptr(dy), ptr(dpre), ctypes.c_longlong(rows), _i(F), cur_stream()))
    return dpre


def project_heads2(x, w, rows_per_img, heads, d, which0, rm=(None, None, None), tr=(None, None, None)):
    """general head-split projection: rm[i] [BH, alloc, dp] row-major, tr[i] [BH, d16, alloc] transposed"""
    M, K = x.shape
    N = w.shape[0]
    VP = ctypes.c_void_p * 3
    IP = ctypes.c_int * 3
    rmp = VP(*[t.data_ptr() if t is not None else None for t in rm])
    trp = VP(*[t.data_ptr() if t is not None else None for t in tr])
    rma = IP(*[t.shape[1] if t is not None else 0 for t in rm])
    tra = IP(*[t.shape[2] if t is not None else 0 for t in tr])
    check(lib().b200lmd_project_heads2_f16(ptr(x), _i(x.stride(0)), ptr(w), _i(M), _i(N), _i(K), _i(rows_per_img),
                                           _i(heads), _i(d), _i(which0), rmp, rma, trp, tra, cur_stream()))


def attention_bwd(q, k, v, dO, qt, kt, dOt, lse2, o_tok, do_tok, B, heads, nq, nk, d, scale, dp_extra=None,
                  want_dkv=True, nk_store=None, use_delta=True):
    C = heads * d
    nk_store = nk if nk_store is None else nk_store
    dq = torch.zeros(B * nq, C, device=q.device, dtype=torch.float16)
    dk = torch.zeros(B * nk_store, C, device=q.device, dtype=torch.float16) if want_dkv else None
    dv = torch.zeros(B * nk_store, C, device=q.device, dtype=torch.float16) if want_dkv else None
    delta = torch.zeros(B * heads, q.shape[1], device=q.device, dtype=torch.float32) if use_delta else None
    check(lib().b200lmd_attention_bwd_f16(
        ptr(q), ptr(k), ptr(v), ptr(dO), ptr(qt), ptr(kt), ptr(dOt), ptr(lse2), ptr(delta), ptr(o_tok),
        _i(o_tok.stride(0) if o_tok is not None else 0), ptr(do_tok),
        _i(do_tok.stride(0) if do_tok is not None else 0), ptr(dp_extra),
        _i(dp_extra.shape[2] if dp_extra is not None else 0), ptr(dq), _i(C), ptr(dk), _i(C), ptr(dv), _i(C),
        _i(nk_store), _i(B), _i(heads), _i(nq), _i(nk), _i(q.shape[1]), _i(k.shape[1]), _i(d), _f(scale),
        cur_stream()))
    return dq, dk, dv


def xattn_fwd(q, k, vt, B, heads, nq, nk, d, scale, loss=None, want_probs=False, save_tok=None, want_lse=False):
    """fused cross-attention (+ guidance loss when `loss` is a guidance.KeyLoss)"""
    out = torch.empty(B * nq, heads * d, device=q.device, dtype=torch.float16)
    lse = torch.zeros(B * heads, q.shape[1], device=q.device, dtype=torch.float32) if want_lse else None
    probs = torch.empty(B * heads, nq, nk, device=q.device, dtype=torch.float16) if want_probs else None
    ptok = torch.zeros(B * heads, nq, device=q.device, dtype=torch.float16) if save_tok is not None else None
    check(lib().b200lmd_xattn_fwd_f16(ptr(q), ptr(k), ptr(vt), ptr(out), _i(heads * d), ptr(lse), ptr(probs),
                                      ptr(save_tok), ptr(ptok), ctypes.byref(loss.c) if loss is not None else None,
                                      _i(B), _i(heads), _i(nq), _i(nk), _i(q.shape[1]), _i(k.shape[1]), _i(d),
                                      _f(scale), cur_stream()))
    return out, lse, probs, ptok


def xattn_fused_supported(heads, d, n):
    return bool(lib().b200lmd_xattn_fused_supported(_i(heads), _i(d), _i(n)))


def xattn_fused(x, wq, k, vt, wo, bias_o, residual, B, n, heads, d, nk, scale, loss=None, want_probs=False,
                save_tok=None, want_q=False):
    """whole cross-attention op (projections + attention + loss) in one launch; x [B*n, C]"""
    C = heads * d
    out = torch.empty(B * n, C, device=x.device, dtype=torch.float16)
    o_scr = torch.empty(B * n, C, device=x.device, dtype=torch.float16)
    q = torch.zeros(B * heads, n, round_dp(d), device=x.device, dtype=torch.float16) if want_q else None
    lse = torch.zeros(B * heads, n, device=x.device, dtype=torch.float32) if want_q else None
    probs = torch.empty(B * heads, n, nk, device=x.device, dtype=torch.float16) if want_probs else None
    ptok = torch.zeros(B * heads, n, device=x.device, dtype=torch.float16) if save_tok is not None else None
    check(lib().b200lmd_xattn_fused_f16(ptr(x), ptr(wq), ptr(k), ptr(vt), ptr(wo), ptr(bias_o), ptr(residual), ptr(out),
                                        ptr(o_scr), ptr(q), ptr(lse), ptr(probs), ptr(save_tok), ptr(ptok),
                                        ctypes.byref(loss.c) if loss is not None else None, _i(B), _i(n), _i(heads),
                                        _i(d), _i(nk), _i(k.shape[1]), _f(scale), cur_stream()))
    return out, q, lse, probs, ptok
