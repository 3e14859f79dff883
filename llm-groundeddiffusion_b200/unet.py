"""Host-side mirror of the reference UNet (models/unet_2d_condition.py:704-980) over hand-written sm_100a kernels.

Python only sequences C-ABI calls (include/b200lmd.h); tensors are torch allocations used as raw device buffers.
Activations are fp16 NHWC ([B, H, W, C] == [B*n, C]), accumulation fp32.  Two entry points:
  forward(...)            full UNet -> eps (CFG pass; optional attention-map saving for the reference's
                          save_attn_to_dict / return_token_ca_only contract, models/attention_processor.py:463-482)
  guidance_gradient(...)  forward truncated at the last guidance key with the fused cross-attention+loss kernel, then
                          a hand-written backward chain (dgrad only - weights are frozen) to d loss / d latent,
                          replacing torch.autograd.grad at models/pipelines.py:56.  No autograd graph is built.
Layer semantics follow the reference files cited per method; weights use diffusers state-dict names.
"""
import ctypes
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

from . import guidance as G
from . import ops
from ._lib import check, cur_stream, lib, ptr

_i, _f = ctypes.c_int, ctypes.c_float


@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    heads: Tuple[int, ...] = (8, 8, 8, 8)
    cross_attention_dim: int = 768
    norm_groups: int = 32
    norm_eps: float = 1e-5
    use_linear_projection: bool = False
    use_gated_attention: bool = False
    down_attn: Tuple[bool, ...] = (True, True, True, False)
    up_attn: Tuple[bool, ...] = (False, True, True, True)

    @staticmethod
    def sd15(gligen=False):
        return UNetConfig(use_gated_attention=gligen)

    @staticmethod
    def sd21():
        return UNetConfig(heads=(5, 10, 20, 20), cross_attention_dim=1024, use_linear_projection=True)

    @staticmethod
    def tiny(gligen=False):
        return UNetConfig(block_out_channels=(128, 256, 512, 512), use_gated_attention=gligen)


class GemmDesc(ctypes.Structure):
    """b200lmd_gemm_desc"""
    _fields_ = [("A", ctypes.c_void_p), ("aB", _i), ("aH", _i), ("aW", _i), ("a_ld", _i), ("Cin", _i),
                ("W", ctypes.c_void_p), ("N", _i), ("wtaps", _i), ("gB", _i), ("gH", _i), ("gW", _i), ("ntaps", _i),
                ("taps", (ctypes.c_short * 4) * 9), ("OH", _i), ("OW", _i), ("sy", _i), ("sx", _i), ("oy", _i),
                ("ox", _i), ("mode", _i), ("alpha", _f), ("bias", ctypes.c_void_p), ("chan_add", ctypes.c_void_p),
                ("rows_per_img", _i), ("residual", ctypes.c_void_p), ("ldr", _i), ("out", ctypes.c_void_p),
                ("ldo", _i), ("out_f32", ctypes.c_void_p), ("ldo32", _i), ("accumulate", _i),
                ("pre", ctypes.c_void_p), ("heads", _i), ("head_dim", _i), ("which0", _i),
                ("rm", ctypes.c_void_p * 3), ("rm_alloc", _i * 3), ("tr", ctypes.c_void_p * 3), ("tr_alloc", _i * 3)]


def _dp(t):
    return t.data_ptr() if t is not None else None


TAPS_3x3 = [(s - 1, r - 1, 0, r * 3 + s) for r in range(3) for s in range(3)]
TAPS_3x3_DGRAD = [(1 - s, 1 - r, 0, r * 3 + s) for r in range(3) for s in range(3)]
TAPS_1x1 = [(0, 0, 0, 0)]


def gemm(A, a_geom, W, N, wtaps, grid, taps, out=None, ldo=0, out_f32=None, ldo32=0, bias=None, chan_add=None,
         rows_per_img=0, residual=None, ldr=0, alpha=1.0, accumulate=False, mode=0, pre=None, omap=None, heads=0,
         head_dim=0, which0=0, rm=(None, None, None), tr=(None, None, None)):
    """one launch of the implicit-GEMM kernel; a_geom = (aB, aH, aW, a_ld, Cin); grid = (gB, gH, gW);
    omap = (OH, OW, sy, sx, oy, ox) or None for the identity pixel mapping"""
    d = GemmDesc()
    d.A = A.data_ptr()
    d.aB, d.aH, d.aW, d.a_ld, d.Cin = a_geom
    d.W = W.data_ptr()
    d.N, d.wtaps = N, wtaps
    d.gB, d.gH, d.gW = grid
    d.ntaps = len(taps)
    for i, t in enumerate(taps):
        for j in range(4):
            d.taps[i][j] = t[j]
    if omap is None:
        omap = (grid[1], grid[2], 1, 1, 0, 0)
    d.OH, d.OW, d.sy, d.sx, d.oy, d.ox = omap
    d.mode, d.alpha = mode, alpha
    d.bias, d.chan_add, d.rows_per_img = _dp(bias), _dp(chan_add), rows_per_img
    d.residual, d.ldr = _dp(residual), ldr
    d.out, d.ldo, d.out_f32, d.ldo32 = _dp(out), ldo, _dp(out_f32), ldo32
    d.accumulate, d.pre = int(accumulate), _dp(pre)
    d.heads, d.head_dim, d.which0 = heads, head_dim, which0
    for i in range(3):
        d.rm[i] = _dp(rm[i])
        d.rm_alloc[i] = rm[i].shape[1] if rm[i] is not None else 0
        d.tr[i] = _dp(tr[i])
        d.tr_alloc[i] = tr[i].shape[2] if tr[i] is not None else 0
    check(lib().b200lmd_gemm(ctypes.byref(d), cur_stream()))


class TextKV:
    """per-prompt cross-attention operands of all transformer layers (timestep independent; computed once per prompt,
    SURVEY.md section 7 step 3): K, V row-major + K^T, V^T slabs for B_text = text.shape[0] samples"""

    def __init__(self):
        self.slabs: Dict[str, tuple] = {}
        self.B = 0
        self.T = 0

    def view(self, prefix, b0, b1, heads):
        k, v, kt, vt = self.slabs[prefix]
        s = slice(b0 * heads, b1 * heads)
        return k[s], v[s], kt[s], vt[s]


class _Truncate(Exception):
    pass


class B200UNet:
    def __init__(self, cfg: UNetConfig, weights: Dict[str, torch.Tensor], device="cuda:0", grad_scale=256.0):
        self.cfg = cfg
        self.config = type("Cfg", (), {"in_channels": cfg.in_channels})()   # utils/latents.py:43,127 reads this
        self.dev = torch.device(device)
        self.gscale = float(grad_scale)
        self.w: Dict[str, torch.Tensor] = {}
        self._prepare(weights)
        self.use_fused_xattn = True
        self._slab_cache: Dict[tuple, torch.Tensor] = {}
        self.tape: Optional[list] = None
        self.grads: Dict[int, torch.Tensor] = {}
        self._keep: list = []

    # ------------------------------------------------------------------------------------------ weight layout
    def _prepare(self, sd):
        cfg, dev = self.cfg, self.dev
        h16 = lambda t: t.to(dev, torch.float16).contiguous()
        f32 = lambda t: t.to(dev, torch.float32).contiguous()
        for name, t in sd.items():
            t = t.detach()
            if name.endswith(".weight") and t.ndim == 4:
                co, ci, kh, kw = t.shape
                base = name[:-7]
                if name == "conv_in.weight":
                    tp = torch.zeros(co, 8, kh, kw, dtype=t.dtype, device=t.device)
                    tp[:, :ci] = t
                    t, ci = tp, 8
                self.w[base + ".w"] = h16(t.permute(0, 2, 3, 1).reshape(co, kh * kw, ci))
                wd = t.permute(1, 2, 3, 0).reshape(ci, kh * kw, co)
                if name == "conv_out.weight":
                    continue  # no gradient ever flows through conv_out
                self.w[base + ".wd"] = h16(wd)
            elif name.endswith(".weight") and t.ndim == 2:
                base = name[:-7]
                if base.endswith("net.0.proj"):
                    F = t.shape[0] // 2
                    w_il, b_il = ops.geglu_interleave(h16(t), f32(sd[base + ".bias"]))
                    self.w[base + ".w_il"], self.w[base + ".b_il"] = w_il, b_il
                    self.w[base + ".wd_il"] = w_il.t().contiguous()
                else:
                    self.w[base + ".w"] = h16(t)
                    self.w[base + ".wd"] = h16(t.t())
            elif name.endswith(".weight") or name.endswith(".bias") or t.ndim <= 1:
                self.w[name] = f32(t)
        # fused projections
        for name in list(sd.keys()):
            if name.endswith("attn1.to_q.weight") or name.endswith("fuser.attn.to_q.weight"):
                p = name[:-len(".to_q.weight")]
                wqkv = torch.cat([sd[p + ".to_q.weight"], sd[p + ".to_k.weight"], sd[p + ".to_v.weight"]], 0)
                self.w[p + ".qkv.w"] = h16(wqkv)
                self.w[p + ".qkv.wd"] = h16(wqkv.t())
            if name.endswith("attn2.to_k.weight"):
                p = name[:-len(".to_k.weight")]
                self.w[p + ".kv.w"] = h16(torch.cat([sd[p + ".to_k.weight"], sd[p + ".to_v.weight"]], 0))
            if name.endswith("fuser.alpha_attn"):
                p = name[:-len(".alpha_attn")]
                ta, td = math.tanh(float(sd[p + ".alpha_attn"])), math.tanh(float(sd[p + ".alpha_dense"]))
                self.w[p + ".tanh_attn"], self.w[p + ".tanh_dense"] = ta, td
                self.w[p + ".attn.to_out.0.bias_g"] = f32(sd[p + ".attn.to_out.0.bias"] * ta)
                self.w[p + ".ff.net.2.bias_g"] = f32(sd[p + ".ff.net.2.bias"] * td)

    # ------------------------------------------------------------------------------------------ autograd-free tape
    def _rec(self, fn):
        if self.tape is not None:
            self.tape.append(fn)

    def _grad_of(self, t):
        return self.grads.get(t.data_ptr())

    def _acc(self, t, like=None):
        """gradient buffer of tensor t and whether it already holds a value (=> accumulate).  Keyed by the data pointer on
        purpose: the conv view [B,H,W,C] and the token view [B*n,C] of one activation are different tensor objects over
        the same memory and must share one gradient; `_keep` pins every keyed tensor for the life of the tape, so a
        pointer cannot be recycled under a live key."""
        g = self.grads.get(t.data_ptr())
        if g is not None:
            return g, True
        g = torch.empty_like(t if like is None else like)
        self.grads[t.data_ptr()] = g
        self._keep.append(t)
        return g, False

    def _add_grad(self, t, g_src):
        """grad(t) += g_src (fp16 [rows, C])"""
        g, acc = self._acc(t)
        C = g.shape[-1]
        rows = g.numel() // C
        check(lib().b200lmd_copy_cols_f16(ptr(g_src), _i(C), _i(0), ptr(g), _i(C), _i(0), ctypes.c_longlong(rows),
                                          _i(C), _i(int(acc)), cur_stream()))

    # ------------------------------------------------------------------------------------------ primitive ops
    def conv(self, x, name, chan_add=None, residual=None, out_f32=False, taps=TAPS_3x3):
        """3x3/1x1 stride-1 conv on NHWC (diffusers ResnetBlock2D.conv1/conv2/conv_shortcut, proj_in/out)"""
        B, H, W, Cin = x.shape
        w = self.w[name + ".w"]
        Cout = w.shape[0]
        y = torch.empty(B, H, W, Cout, device=self.dev, dtype=torch.float16) if not out_f32 else None
        y32 = torch.empty(B, H, W, Cout, device=self.dev, dtype=torch.float32) if out_f32 else None
        gemm(x, (B, H, W, Cin, Cin), w, Cout, w.shape[1], (B, H, W), taps, out=y, ldo=Cout, out_f32=y32, ldo32=Cout,
             bias=self.w.get(name + ".bias"), chan_add=chan_add, rows_per_img=H * W, residual=residual, ldr=Cout)
        out = y32 if out_f32 else y
        if self.tape is not None:
            def bwd():
                dy = self._grad_of(out)
                if dy is None:
                    return
                if residual is not None:
                    self._add_grad(residual, dy)
                wd = self.w[name + ".wd"]
                if name == "conv_in":      # end of the chain: fp32 d loss / d latent (NHWC-8), still times gscale
                    self.latent_grad = torch.empty(B, H * W, Cin, device=self.dev, dtype=torch.float32)
                    gemm(dy, (B, H, W, Cout, Cout), wd, Cin, 9, (B, H, W), TAPS_3x3_DGRAD, out_f32=self.latent_grad,
                         ldo32=Cin)
                    return
                dx, acc = self._acc(x)
                gemm(dy, (B, H, W, Cout, Cout), wd, Cin, wd.shape[1], (B, H, W),
                     TAPS_3x3_DGRAD if len(taps) == 9 else TAPS_1x1, out=dx, ldo=Cin, accumulate=acc)
            self._rec(bwd)
        return out

    def conv_down(self, x, name):
        """Downsample2D: conv3x3 stride 2 pad 1, via a parity space-to-depth copy"""
        B, H, W, C = x.shape
        Ho, Wo = H // 2, W // 2
        s2d = torch.empty(4 * B, Ho, Wo, C, device=self.dev, dtype=torch.float16)
        check(lib().b200lmd_space_to_depth_f16(ptr(x), ptr(s2d), _i(B), _i(Ho), _i(Wo), _i(C), cur_stream()))
        w = self.w[name + ".w"]
        Cout = w.shape[0]
        taps = []
        for r in range(3):
            py, dy = (1, -1) if r == 0 else ((0, 0) if r == 1 else (1, 0))
            for s in range(3):
                px, dx = (1, -1) if s == 0 else ((0, 0) if s == 1 else (1, 0))
                taps.append((dx, dy, (py * 2 + px) * B, r * 3 + s))
        y = torch.empty(B, Ho, Wo, Cout, device=self.dev, dtype=torch.float16)
        gemm(s2d, (4 * B, Ho, Wo, C, C), w, Cout, 9, (B, Ho, Wo), taps, out=y, ldo=Cout, bias=self.w[name + ".bias"])
        if self.tape is not None:
            def bwd():
                dyv = self._grad_of(y)
                if dyv is None:
                    return
                dx, acc = self._acc(x)
                wd = self.w[name + ".wd"]
                for PY in range(2):
                    rs = [(1, 0)] if PY == 0 else [(0, 1), (2, 0)]      # (r, dy offset into the half-res grid)
                    for PX in range(2):
                        ss = [(1, 0)] if PX == 0 else [(0, 1), (2, 0)]
                        tp = [(dxo, dyo, 0, r * 3 + s) for (r, dyo) in rs for (s, dxo) in ss]
                        gemm(dyv, (B, Ho, Wo, Cout, Cout), wd, C, 9, (B, Ho, Wo), tp, out=dx, ldo=C, accumulate=acc,
                             omap=(H, W, 2, 2, PY, PX))
            self._rec(bwd)
        return y

    def upsample_conv(self, x, name):
        """Upsample2D: nearest x2 then conv3x3"""
        B, H, W, C = x.shape
        up = torch.empty(B, 2 * H, 2 * W, C, device=self.dev, dtype=torch.float16)
        check(lib().b200lmd_upsample2x_f16(ptr(x), ptr(up), _i(B), _i(H), _i(W), _i(C), cur_stream()))
        if self.tape is not None:
            def bwd():
                d = self._grad_of(up)
                if d is None:
                    return
                dx, acc = self._acc(x)
                check(lib().b200lmd_upsample2x_bwd_f16(ptr(d), ptr(dx), _i(B), _i(H), _i(W), _i(C), _i(int(acc)),
                                                       cur_stream()))
            self._rec(bwd)
        return self.conv(up, name)

    def concat(self, a, b):
        B, H, W, Ca = a.shape
        Cb = b.shape[3]
        y = torch.empty(B, H, W, Ca + Cb, device=self.dev, dtype=torch.float16)
        rows = B * H * W
        L = ctypes.c_longlong(rows)
        check(lib().b200lmd_copy_cols_f16(ptr(a), _i(Ca), _i(0), ptr(y), _i(Ca + Cb), _i(0), L, _i(Ca), _i(0), cur_stream()))
        check(lib().b200lmd_copy_cols_f16(ptr(b), _i(Cb), _i(0), ptr(y), _i(Ca + Cb), _i(Ca), L, _i(Cb), _i(0), cur_stream()))
        if self.tape is not None:
            def bwd():
                d = self._grad_of(y)
                if d is None:
                    return
                for t, off, Ct in ((a, 0, Ca), (b, Ca, Cb)):
                    g, acc = self._acc(t)
                    check(lib().b200lmd_copy_cols_f16(ptr(d), _i(Ca + Cb), _i(off), ptr(g), _i(Ct), _i(0), L, _i(Ct),
                                                      _i(int(acc)), cur_stream()))
            self._rec(bwd)
        return y

    def group_norm(self, x, name, eps, silu):
        B, H, W, C = x.shape
        g, bta = self.w[name + ".weight"], self.w[name + ".bias"]
        y, sums = ops.groupnorm(x.view(B, H * W, C), g, bta, self.cfg.norm_groups, eps, silu, want_sums=True)
        y = y.view(B, H, W, C)
        if self.tape is not None:
            def bwd():
                d = self._grad_of(y)
                if d is None:
                    return
                dx, acc = self._acc(x)
                bsums = torch.empty_like(sums)
                check(lib().b200lmd_groupnorm_bwd_f16(ptr(d), ptr(x), ptr(sums), ptr(g), ptr(bta), ptr(dx), ptr(bsums),
                                                      _i(B), _i(H * W), _i(C), _i(self.cfg.norm_groups), _f(eps),
                                                      _i(int(silu)), _i(int(acc)), cur_stream()))
            self._rec(bwd)
        return y

    def layer_norm(self, x, name):
        g, bta = self.w[name + ".weight"], self.w[name + ".bias"]
        rec = self.tape is not None
        r = ops.layernorm(x, g, bta, want_stats=rec)
        if not rec:
            return r
        y, stats = r

        def bwd():
            d = self._grad_of(y)
            if d is None:
                return
            dx, acc = self._acc(x)
            check(lib().b200lmd_layernorm_bwd_f16(ptr(d), ptr(x), ptr(stats), ptr(g), ptr(dx),
                                                  ctypes.c_longlong(x.shape[0]), _i(x.shape[1]), _i(int(acc)),
                                                  cur_stream()))
        self._rec(bwd)
        return y

    def linear(self, x, name, residual=None, alpha=1.0, bias=None, out_f32=False):
        """x [M,K] -> [M,N] (+bias)(+residual); bias defaults to <name>.bias when present"""
        M, K = x.shape
        w = self.w[name + ".w"]
        N = w.shape[0]
        if bias is None:
            bias = self.w.get(name + ".bias")
        y = torch.empty(M, N, device=self.dev, dtype=torch.float32 if out_f32 else torch.float16)
        gemm(x, (1, 1, M, x.stride(0), K), w, N, 1, (1, 1, M), TAPS_1x1, out=None if out_f32 else y, ldo=N,
             out_f32=y if out_f32 else None, ldo32=N, bias=bias, residual=residual, ldr=N, alpha=alpha)
        if self.tape is not None:
            def bwd():
                d = self._grad_of(y)
                if d is None:
                    return
                if residual is not None:
                    self._add_grad(residual, d)
                dx, acc = self._acc(x)
                wd = self.w[name + ".wd"]
                gemm(d, (1, 1, M, N, N), wd, K, 1, (1, 1, M), TAPS_1x1, out=dx, ldo=K, alpha=alpha, accumulate=acc)
            self._rec(bwd)
        return y

    def feed_forward(self, xn, prefix, residual, alpha=1.0, out_bias=None):
        """FeedForward(GEGLU) (models/attention.py:286-335): net.0.proj (GEGLU fused in the GEMM epilogue) -> net.2"""
        M, K = xn.shape
        w_il, b_il = self.w[prefix + ".net.0.proj.w_il"], self.w[prefix + ".net.0.proj.b_il"]
        F = w_il.shape[0] // 2
        rec = self.tape is not None
        h = torch.empty(M, F, device=self.dev, dtype=torch.float16)
        pre = torch.empty(M, 2 * F, device=self.dev, dtype=torch.float16) if rec else None
        gemm(xn, (1, 1, M, K, K), w_il, 2 * F, 1, (1, 1, M), TAPS_1x1, out=h, ldo=F, bias=b_il, mode=1, pre=pre)
        if rec:
            def bwd():
                d = self._grad_of(h)
                if d is None:
                    return
                dpre = ops.geglu_bwd(pre, d)
                dx, acc = self._acc(xn)
                wd = self.w[prefix + ".net.0.proj.wd_il"]
                gemm(dpre, (1, 1, M, 2 * F, 2 * F), wd, K, 1, (1, 1, M), TAPS_1x1, out=dx, ldo=K, accumulate=acc)
            self._rec(bwd)
        return self.linear(h, prefix + ".net.2", residual=residual, alpha=alpha, bias=out_bias)

    # ------------------------------------------------------------------------------------------ attention
    def _slabs(self, BH, n_alloc, d, rm=True, tr=True, slot=None, owner=None):
        """attention operand slabs.  Padding (columns d..dp, rows d..d16) must be zero and is never written, so the
        zero-filled buffers are cached and reused instead of memset-ing them per call (~1 GB per attention layer at
        batch 64): in forward-only mode one buffer per (shape, slot) serves every layer - stream order serialises the
        layers; with the tape on, the operands must survive until the layer's backward, so the cache is per layer
        (`owner`) as well."""
        dp, d16 = ops.round_dp(d), ops.round_d16(d)

        def get(kind, shape):
            if slot is None or (self.tape is not None and owner is None):
                return torch.zeros(shape, device=self.dev, dtype=torch.float16)
            key = (kind, slot, owner if self.tape is not None else None) + tuple(shape)
            t = self._slab_cache.get(key)
            if t is None:
                t = self._slab_cache[key] = torch.zeros(shape, device=self.dev, dtype=torch.float16)
            return t
        a = get("rm", (BH, n_alloc, dp)) if rm else None
        b = get("tr", (BH, d16, n_alloc)) if tr else None
        return a, b

    def self_attention(self, xn, B, n, prefix, heads, residual, alpha=1.0, out_bias=None, nk_store=None):
        """attn1 / fuser.attn: q,k,v from the same tokens (models/attention_processor.py:305-375)"""
        C = xn.shape[1]
        d = C // heads
        rec = self.tape is not None
        na = (n + 7) // 8 * 8
        q, qt = self._slabs(B * heads, na, d, True, rec, slot="q", owner=prefix)
        k, kt = self._slabs(B * heads, na, d, True, rec, slot="k", owner=prefix)
        v, vt = self._slabs(B * heads, na, d, rec, True, slot="v", owner=prefix)
        M = B * n
        gemm(xn, (1, 1, M, C, C), self.w[prefix + ".qkv.w"], 3 * C, 1, (1, 1, M), TAPS_1x1, mode=2, rows_per_img=n,
             heads=heads, head_dim=d, which0=0, rm=(q, k, v), tr=(qt, kt, vt))
        scale = d ** -0.5
        o = torch.empty(M, C, device=self.dev, dtype=torch.float16)
        lse = torch.zeros(B * heads, na, device=self.dev, dtype=torch.float32) if rec else None
        check(lib().b200lmd_attention_fwd_f16(ptr(q), ptr(k), ptr(vt), ptr(o), _i(C), ptr(lse), _i(B), _i(heads), _i(n),
                                              _i(n), _i(na), _i(na), _i(d), _f(scale), cur_stream()))
        y = self.linear(o, prefix + ".to_out.0", residual=residual, alpha=alpha, bias=out_bias)
        if rec:
            self.tape.pop()          # replace the generic linear backward by one that lands in head slabs

            def bwd():
                dyv = self._grad_of(y)
                if dyv is None:
                    return
                if residual is not None:
                    self._add_grad(residual, dyv)
                dO, dOt = self._slabs(B * heads, na, d, True, True, slot="dO", owner=prefix)
                gemm(dyv, (1, 1, M, C, C), self.w[prefix + ".to_out.0.wd"], C, 1, (1, 1, M), TAPS_1x1, mode=2,
                     rows_per_img=n, heads=heads, head_dim=d, which0=0, rm=(dO, None, None), tr=(dOt, None, None),
                     alpha=alpha)
                delta = torch.zeros(B * heads, na, device=self.dev, dtype=torch.float32)
                check(lib().b200lmd_attn_delta_slab(ptr(dO), ptr(o), _i(C), ptr(delta), _i(B), _i(heads), _i(n), _i(na),
                                                    _i(d), cur_stream()))
                ns = n if nk_store is None else nk_store
                dqkv = torch.zeros(M, 3 * C, device=self.dev, dtype=torch.float16)
                if ns != n:
                    raise NotImplementedError
                check(lib().b200lmd_attention_bwd_f16(
                    ptr(q), ptr(k), ptr(v), ptr(dO), ptr(qt), ptr(kt), ptr(dOt), ptr(lse), ptr(delta), None, _i(0),
                    None, _i(0), None, _i(0), ptr(dqkv), _i(3 * C), ctypes.c_void_p(dqkv.data_ptr() + 2 * C),
                    _i(3 * C), ctypes.c_void_p(dqkv.data_ptr() + 4 * C), _i(3 * C), _i(n), _i(B), _i(heads), _i(n),
                    _i(n), _i(na), _i(na), _i(d), _f(scale), cur_stream()))
                dx, acc = self._acc(xn)
                gemm(dqkv, (1, 1, M, 3 * C, 3 * C), self.w[prefix + ".qkv.wd"], C, 1, (1, 1, M), TAPS_1x1, out=dx,
                     ldo=C, accumulate=acc)
            self._rec(bwd)
        return y

    def cross_attention(self, xn, B, n, prefix, heads, key, residual, kv, loss=None, save=None):
        """attn2 through the fused cross-attention(+loss) kernel (models/attention_processor.py:407-483 and
        utils/guidance.py:244-286); kv = (k, v, kt, vt) slabs of the B text rows; save: dict(probs=bool, tok=int32[B])"""
        C = xn.shape[1]
        d = C // heads
        rec = self.tape is not None
        na = (n + 7) // 8 * 8
        M = B * n
        k, v, kt, vt = kv
        T = self.text_T
        scale = d ** -0.5
        want_probs = bool(save and save.get("probs"))
        tok = save.get("tok") if save else None
        probs = torch.empty(B * heads, n, T, device=self.dev, dtype=torch.float16) if want_probs else None
        ptok = torch.zeros(B * heads, n, device=self.dev, dtype=torch.float16) if tok is not None else None
        lse = torch.zeros(B * heads, na, device=self.dev, dtype=torch.float32) if rec else None
        fused = self.use_fused_xattn and T <= 80 and k.shape[1] >= 80 and ops.xattn_fused_supported(heads, d, n)
        # loss.c is None for an external-gradient holder (adapter.py): d loss / dP arrives through loss.dp_extra
        loss_c = ctypes.byref(loss.c) if (loss is not None and getattr(loss, "c", None) is not None) else None
        if fused:
            # whole op in ONE launch (csrc/xattn_fused.cuh): to_q, QK^T, softmax, loss, PV, to_out + bias + residual
            q = torch.zeros(B * heads, na, ops.round_dp(d), device=self.dev, dtype=torch.float16) if rec else None
            y = torch.empty(M, C, device=self.dev, dtype=torch.float16)
            o = torch.empty(M, C, device=self.dev, dtype=torch.float16)
            check(lib().b200lmd_xattn_fused_f16(
                ptr(xn), ptr(self.w[prefix + ".to_q.w"]), ptr(k), ptr(vt), ptr(self.w[prefix + ".to_out.0.w"]),
                ptr(self.w[prefix + ".to_out.0.bias"]), ptr(residual), ptr(y), ptr(o), ptr(q), ptr(lse), ptr(probs),
                ptr(tok), ptr(ptok), loss_c, _i(B), _i(n), _i(heads),
                _i(d), _i(T), _i(k.shape[1]), _f(scale), cur_stream()))
        else:
            q, _ = self._slabs(B * heads, na, d, True, False, slot="xq", owner=prefix)
            gemm(xn, (1, 1, M, C, C), self.w[prefix + ".to_q.w"], C, 1, (1, 1, M), TAPS_1x1, mode=2, rows_per_img=n,
                 heads=heads, head_dim=d, which0=0, rm=(q, None, None))
            o = torch.empty(M, C, device=self.dev, dtype=torch.float16)
            check(lib().b200lmd_xattn_fwd_f16(ptr(q), ptr(k), ptr(vt), ptr(o), _i(C), ptr(lse), ptr(probs), ptr(tok),
                                              ptr(ptok), loss_c, _i(B),
                                              _i(heads), _i(n), _i(T), _i(na), _i(k.shape[1]), _i(d), _f(scale),
                                              cur_stream()))
        if save is not None:
            save["out"][key] = dict(probs=probs.view(B, heads, n, T) if probs is not None else None,
                                    tok=ptok.view(B, heads, n) if ptok is not None else None)
        if loss is not None and key == self._last_key:
            # nothing after the last guidance map can influence d loss / d latent: stop the forward here
            if rec:
                self._rec(lambda: self._xattn_bwd(xn, q, k, v, kt, None, lse, loss, B, heads, n, na, d, scale, prefix,
                                                  C, M))
            raise _Truncate()
        if not fused:
            y = self.linear(o, prefix + ".to_out.0", residual=residual)
            if rec:
                self.tape.pop()
        if rec:
            def bwd():
                dyv = self._grad_of(y)
                if dyv is None and loss is None:
                    return
                dO = None
                if dyv is not None:
                    if residual is not None:
                        self._add_grad(residual, dyv)
                    dO, _ = self._slabs(B * heads, na, d, True, False, slot="xdO", owner=prefix)
                    gemm(dyv, (1, 1, M, C, C), self.w[prefix + ".to_out.0.wd"], C, 1, (1, 1, M), TAPS_1x1, mode=2,
                         rows_per_img=n, heads=heads, head_dim=d, which0=0, rm=(dO, None, None))
                self._xattn_bwd(xn, q, k, v, kt, dO, lse, loss, B, heads, n, na, d, scale, prefix, C, M)
            self._rec(bwd)
        return y

    def _xattn_bwd(self, xn, q, k, v, kt, dO, lse, loss, B, heads, n, na, d, scale, prefix, C, M):
        dq = torch.zeros(M, C, device=self.dev, dtype=torch.float16)
        ext = loss.dp_extra if loss is not None else None
        check(lib().b200lmd_attention_bwd_f16(
            ptr(q), ptr(k), ptr(v), ptr(dO), None, ptr(kt), None, ptr(lse), None, None, _i(0), None, _i(0), ptr(ext),
            _i(ext.shape[2] if ext is not None else 0), ptr(dq), _i(C), None, _i(0), None, _i(0), _i(0), _i(B),
            _i(heads), _i(n), _i(self.text_T), _i(na), _i(k.shape[1]), _i(d), _f(scale), cur_stream()))
        dx, acc = self._acc(xn)
        gemm(dq, (1, 1, M, C, C), self.w[prefix + ".to_q.wd"], C, 1, (1, 1, M), TAPS_1x1, out=dx, ldo=C, accumulate=acc)

    # ------------------------------------------------------------------------------------------ text / time
    def set_text(self, text, kv=None):
        """text [Bt, T, ctx] -> K / V slabs of every attn2 layer (once per prompt; K/V do not depend on t or z).
        kv: a TextKV of the same shape to refill in place (its slab addresses are baked into captured CUDA graphs)"""
        Bt, T, ctx = text.shape
        x = text.to(self.dev, torch.float16).reshape(Bt * T, ctx).contiguous()
        reuse = kv is not None and kv.B == Bt and kv.T == T
        if not reuse:
            kv = TextKV()
            kv.B, kv.T = Bt, T
        self.text_T = T
        Ta = (T + 7) // 8 * 8
        for p, heads in self._attn_layers():
            w = self.w[p + ".attn2.kv.w"]
            C = w.shape[0] // 2
            d = C // heads
            if reuse:
                k, v, kt, vt = kv.slabs[p]
            else:
                k, kt = self._slabs(Bt * heads, Ta, d)
                v, vt = self._slabs(Bt * heads, Ta, d)
            gemm(x, (1, 1, Bt * T, ctx, ctx), w, 2 * C, 1, (1, 1, Bt * T), TAPS_1x1, mode=2, rows_per_img=T,
                 heads=heads, head_dim=d, which0=1, rm=(None, k, v), tr=(None, kt, vt))
            kv.slabs[p] = (k, v, kt, vt)
        return kv

    def _attn_layers(self):
        cfg = self.cfg
        nb = len(cfg.block_out_channels)
        out = []
        for i in range(nb):
            if cfg.down_attn[i]:
                for j in range(cfg.layers_per_block):
                    out.append((f"down_blocks.{i}.attentions.{j}.transformer_blocks.0", cfg.heads[i]))
        out.append(("mid_block.attentions.0.transformer_blocks.0", cfg.heads[-1]))
        rh = list(reversed(cfg.heads))
        for i in range(nb):
            if cfg.up_attn[i]:
                for j in range(cfg.layers_per_block + 1):
                    out.append((f"up_blocks.{i}.attentions.{j}.transformer_blocks.0", rh[i]))
        return out

    def _time_embedding(self, t):
        """Timesteps(flip_sin_to_cos, shift 0) -> linear_1 -> SiLU -> linear_2 -> SiLU (the resnets apply SiLU to temb
        before time_emb_proj, diffusers ResnetBlock2D); returns fp16 [B, 4*C0] = silu(temb)"""
        tape, self.tape = self.tape, None
        B = t.shape[0]
        c0 = self.cfg.block_out_channels[0]
        e = torch.empty(B, c0, device=self.dev, dtype=torch.float16)
        check(lib().b200lmd_timestep_embed(ptr(t), ptr(e), _i(B), _i(c0), cur_stream()))
        h = self.linear(e, "time_embedding.linear_1", out_f32=True)
        h16 = torch.empty(h.shape, device=self.dev, dtype=torch.float16)
        check(lib().b200lmd_silu_f32_to_f16(ptr(h), ptr(h16), ctypes.c_longlong(h.numel()), cur_stream()))
        h = self.linear(h16, "time_embedding.linear_2", out_f32=True)
        out = torch.empty(h.shape, device=self.dev, dtype=torch.float16)
        check(lib().b200lmd_silu_f32_to_f16(ptr(h), ptr(out), ctypes.c_longlong(h.numel()), cur_stream()))
        self.tape = tape
        return out

    # ------------------------------------------------------------------------------------------ blocks
    def resnet(self, x, temb_act, prefix):
        """diffusers 0.18 ResnetBlock2D (time_embedding_norm='default', output_scale_factor 1)"""
        eps = self.cfg.norm_eps
        h = self.group_norm(x, prefix + ".norm1", eps, True)
        tape, self.tape = self.tape, None
        tproj = self.linear(temb_act, prefix + ".time_emb_proj", out_f32=True)
        self.tape = tape
        h = self.conv(h, prefix + ".conv1", chan_add=tproj)
        h = self.group_norm(h, prefix + ".norm2", eps, True)
        sc = x
        if (prefix + ".conv_shortcut.w") in self.w:
            sc = self.conv(x, prefix + ".conv_shortcut", taps=TAPS_1x1)
        return self.conv(h, prefix + ".conv2", residual=sc)

    def transformer(self, x, prefix, heads, key, st):
        """Transformer2DModel + BasicTransformerBlock (models/transformer_2d.py:216-367, models/attention.py:156-237)"""
        B, H, W, C = x.shape
        n = H * W
        h = self.group_norm(x, prefix + ".norm", 1e-6, False)
        if not self.cfg.use_linear_projection:
            tok = self.conv(h, prefix + ".proj_in", taps=TAPS_1x1).view(B * n, C)
        else:
            tok = self.linear(h.view(B * n, C), prefix + ".proj_in")
        b = prefix + ".transformer_blocks.0"
        tok = self.self_attention(self.layer_norm(tok, b + ".norm1"), B, n, b + ".attn1", heads, residual=tok)
        if self.cfg.use_gated_attention and st.get("objs") is not None and st.get("fuser_on"):
            tok = self.fuser(tok, B, n, b + ".fuser", heads, st["objs"])
        loss = st["loss"].get(key) if st.get("loss") else None
        save = None
        if st.get("save") is not None and (st["save"]["keys"] is None or key in st["save"]["keys"]):
            save = st["save"]
        tok = self.cross_attention(self.layer_norm(tok, b + ".norm2"), B, n, b + ".attn2", heads, key, tok,
                                   st["kv"](b), loss=loss, save=save)
        tok = self.feed_forward(self.layer_norm(tok, b + ".norm3"), b + ".ff", residual=tok)
        if not self.cfg.use_linear_projection:
            return self.conv(tok.view(B, H, W, C), prefix + ".proj_out", residual=x, taps=TAPS_1x1)
        return self.linear(tok, prefix + ".proj_out", residual=x.view(B * n, C)).view(B, H, W, C)

    def fuser(self, tok, B, n, prefix, heads, objs):
        """GatedSelfAttentionDense (models/attention.py:43-53): x += tanh(a)*SelfAttn(LN([x; W objs]))[:, :n];
        x += tanh(b)*FF(LN(x)).  objs [B*30, ctx] fp16"""
        C = tok.shape[1]
        n_obj = objs.shape[0] // B
        nt = n + n_obj
        tape, self.tape = self.tape, None
        o = self.linear(objs, prefix + ".linear")                       # constants w.r.t. the latent
        self.tape = tape
        cat = torch.empty(B * nt, C, device=self.dev, dtype=torch.float16)
        S = cur_stream()
        check(lib().b200lmd_copy_rows_f16(ptr(tok), _i(n), _i(0), ptr(cat), _i(nt), _i(0), _i(B), _i(n), _i(C), _f(1.0),
                                          _i(0), S))
        check(lib().b200lmd_copy_rows_f16(ptr(o), _i(n_obj), _i(0), ptr(cat), _i(nt), _i(n), _i(B), _i(n_obj), _i(C),
                                          _f(1.0), _i(0), S))
        if self.tape is not None:
            def bwd_cat():
                d = self._grad_of(cat)
                if d is None:
                    return
                g, acc = self._acc(tok)
                check(lib().b200lmd_copy_rows_f16(ptr(d), _i(nt), _i(0), ptr(g), _i(n), _i(0), _i(B), _i(n), _i(C),
                                                  _f(1.0), _i(int(acc)), cur_stream()))
            self._rec(bwd_cat)
        ta, td = self.w[prefix + ".tanh_attn"], self.w[prefix + ".tanh_dense"]
        a = self.self_attention(self.layer_norm(cat, prefix + ".norm1"), B, nt, prefix + ".attn", heads, residual=None,
                                alpha=ta, out_bias=self.w[prefix + ".attn.to_out.0.bias_g"])
        x2 = torch.empty_like(tok)
        check(lib().b200lmd_copy_rows_f16(ptr(tok), _i(n), _i(0), ptr(x2), _i(n), _i(0), _i(B), _i(n), _i(C), _f(1.0),
                                          _i(0), S))
        check(lib().b200lmd_copy_rows_f16(ptr(a), _i(nt), _i(0), ptr(x2), _i(n), _i(0), _i(B), _i(n), _i(C), _f(1.0),
                                          _i(1), S))
        if self.tape is not None:
            def bwd_gate():
                d = self._grad_of(x2)
                if d is None:
                    return
                self._add_grad(tok, d)
                g, acc = self._acc(a)
                if not acc:
                    g.zero_()
                check(lib().b200lmd_copy_rows_f16(ptr(d), _i(n), _i(0), ptr(g), _i(nt), _i(0), _i(B), _i(n), _i(C),
                                                  _f(1.0), _i(1), cur_stream()))
            self._rec(bwd_gate)
        return self.feed_forward(self.layer_norm(x2, prefix + ".norm2"), prefix + ".ff", residual=x2, alpha=td,
                                 out_bias=self.w[prefix + ".ff.net.2.bias_g"])

    # ------------------------------------------------------------------------------------------ whole network
    def _network(self, z, t, rep, st):
        """z fp32 [Bz, Cz, H, W]; batch seen by the network = Bz*rep ([uncond copies ; cond copies] for rep=2)"""
        cfg = self.cfg
        Bz, Cz, H, W = z.shape
        B = Bz * rep
        x0 = torch.empty(B, H, W, 8, device=self.dev, dtype=torch.float16)
        check(lib().b200lmd_pack_latents(ptr(z), ptr(x0), _i(Bz), _i(Cz), _i(H * W), _i(rep), cur_stream()))
        temb = self._time_embedding(t)
        h = self.conv(x0, "conv_in")
        skips = [h]
        nb = len(cfg.block_out_channels)
        for i in range(nb):
            for j in range(cfg.layers_per_block):
                h = self.resnet(h, temb, f"down_blocks.{i}.resnets.{j}")
                if cfg.down_attn[i]:
                    h = self.transformer(h, f"down_blocks.{i}.attentions.{j}", cfg.heads[i], ("down", i, j, 0), st)
                skips.append(h)
            if i < nb - 1:
                h = self.conv_down(h, f"down_blocks.{i}.downsamplers.0.conv")
                skips.append(h)
        h = self.resnet(h, temb, "mid_block.resnets.0")
        h = self.transformer(h, "mid_block.attentions.0", cfg.heads[-1], ("mid", 0, 0, 0), st)
        h = self.resnet(h, temb, "mid_block.resnets.1")
        rheads = list(reversed(cfg.heads))
        for i in range(nb):
            for j in range(cfg.layers_per_block + 1):
                h = self.concat(h, skips.pop())
                h = self.resnet(h, temb, f"up_blocks.{i}.resnets.{j}")
                if cfg.up_attn[i]:
                    h = self.transformer(h, f"up_blocks.{i}.attentions.{j}", rheads[i], ("up", i, j, 0), st)
            if i < nb - 1:
                h = self.upsample_conv(h, f"up_blocks.{i}.upsamplers.0.conv")
        h = self.group_norm(h, "conv_norm_out", cfg.norm_eps, True)
        return self.conv(h, "conv_out", out_f32=True)        # fp32 NHWC [B, H, W, 4]

    def forward(self, z, t, kv: TextKV, rep=2, objs=None, fuser_on=False, save_keys=None, save_tok=None,
                save_probs=False):
        """CFG pass.  z fp32 [Bz,4,H,W]; t fp32 [Bz*rep]; kv holds Bz*rep text rows.  Returns (eps fp32 NHWC
        [Bz*rep,H,W,4], saved maps dict key -> {probs [B,heads,n,T] | tok [B,heads,n]})"""
        self.tape = None
        saved = {}
        st = dict(kv=lambda b: kv.slabs[b], objs=objs, fuser_on=fuser_on, loss=None, save=None)
        if save_keys is not None or save_tok is not None or save_probs:
            st["save"] = dict(keys=save_keys, probs=save_probs, tok=save_tok, out=saved)
        self._last_key = None
        eps = self._network(z, t, rep, st)
        return eps, saved

    def guidance_forward(self, z, t, kv_cond, losses, objs=None, fuser_on=False, save=None):
        """cond-only pass truncated at the last guidance key with the hand-written tape recording; returns the tape"""
        self.tape, self.grads, self._keep = [], {}, []
        self.latent_grad = None
        order = [k for _, k in self._key_order() if k in losses]
        self._last_key = order[-1]
        st = dict(kv=kv_cond, objs=objs, fuser_on=fuser_on, loss=losses, save=save)
        try:
            self._network(z, t, 1, st)
            raise RuntimeError("guidance keys never reached")
        except _Truncate:
            pass
        tape, self.tape = self.tape, None
        return tape, order

    def guidance_backward(self, tape):
        """replay the tape backwards (dgrad only - weights are frozen) down to d loss / d latent (fp32 NHWC-8, x gscale)"""
        for fn in reversed(tape):
            fn()
        g = self.latent_grad
        self.grads, self._keep = {}, []
        return g

    def guidance_gradient_launch(self, z, t, kv_cond, losses: Dict[tuple, "G.KeyLoss"], objs=None, fuser_on=False):
        """launch-only part (CUDA-graph capturable: no host synchronisation): cond-only pass truncated at the last
        guidance key + hand-written backward.  Returns (grad fp32 NHWC-8 [B, HW, 8] = gscale * d(loss*loss_scale)/dz,
        loss partials [n_keys, B*heads] on the device)"""
        tape, order = self.guidance_forward(z, t, kv_cond, losses, objs=objs, fuser_on=fuser_on)
        g = self.guidance_backward(tape)
        parts = torch.stack([losses[k].loss_part for k in order])
        return g, parts

    @staticmethod
    def reduce_loss(parts, B):
        """per-image scaled loss: fixed-order host sum of the per-(key, image, head) partials the kernels wrote"""
        p = parts.cpu().numpy()
        return p.reshape(p.shape[0], B, -1).sum(axis=2).sum(axis=0)

    def guidance_gradient(self, z, t, kv_cond, losses, objs=None, fuser_on=False):
        g, parts = self.guidance_gradient_launch(z, t, kv_cond, losses, objs=objs, fuser_on=fuser_on)
        return g, self.reduce_loss(parts, z.shape[0])

    def _key_order(self):
        cfg = self.cfg
        nb = len(cfg.block_out_channels)
        out = []
        for i in range(nb):
            if cfg.down_attn[i]:
                for j in range(cfg.layers_per_block):
                    out.append((f"down_blocks.{i}.attentions.{j}", ("down", i, j, 0)))
        out.append(("mid_block.attentions.0", ("mid", 0, 0, 0)))
        for i in range(nb):
            if cfg.up_attn[i]:
                for j in range(cfg.layers_per_block + 1):
                    out.append((f"up_blocks.{i}.attentions.{j}", ("up", i, j, 0)))
        return out

    def position_net(self, boxes, masks, emb):
        """PositionNet (models/unet_2d_condition.py:79-114): boxes [B,N,4], masks [B,N], emb [B,N,768] -> objs fp16
        [B*N, cross_attention_dim].  Runs once per denoising loop (independent of t and z)."""
        tape, self.tape = self.tape, None
        B, N, D = emb.shape
        rows = B * N
        f = lambda x: x.to(self.dev, torch.float32).contiguous()
        x = torch.empty(rows, D + 64, device=self.dev, dtype=torch.float16)
        bx, mk, em = f(boxes), f(masks), f(emb)      # locals keep the converted copies alive across the launch
        check(lib().b200lmd_position_embed(ptr(bx), ptr(mk), ptr(em),
                                           ptr(self.w["position_net.null_positive_feature"]),
                                           ptr(self.w["position_net.null_position_feature"]), ptr(x), _i(rows), _i(D),
                                           cur_stream()))
        for li in ("0", "2"):
            h = self.linear(x, "position_net.linears." + li, out_f32=True)
            x = torch.empty(h.shape, device=self.dev, dtype=torch.float16)
            check(lib().b200lmd_silu_f32_to_f16(ptr(h), ptr(x), ctypes.c_longlong(h.numel()), cur_stream()))
        out = self.linear(x, "position_net.linears.4")
        self.tape = tape
        return out
