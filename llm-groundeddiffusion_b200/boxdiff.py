"""BoxDiff guidance - host side of csrc/boxdiff.cuh.

Mirrors utils/boxdiff.py of the reference: compute_ca_loss_boxdiff (:120-161), _compute_max_attention_per_index
(:20-101), _compute_loss (:104-117) and the update rule of latent_backward_guidance_boxdiff (:190-259).  This module only
builds the integer tables (cell masks, corner masks, k_fg / k_bg - all of which must equal the reference's bit for bit)
and sequences three launches per guidance step:
    truncated UNet forward saving the fp16 maps of the guidance keys   (B200UNet.guidance_forward)
    b200lmd_boxdiff_loss: loss per image + d loss / dP into every key's dP_extra
    hand-written backward chain to d loss / d latent                     (B200UNet.guidance_backward)
Deviation (stated in DESIGN.md): a phrase whose box covers fewer than 1/P cells gives k = 0 and the reference then
takes the mean of an empty top-k (NaN loss, NaN image); here that raises ValueError when the tables are built.
"""
import ctypes
import math
from dataclasses import dataclass, field
from typing import List

import numpy as np
import torch

from . import guidance as G
from ._lib import check, cur_stream, lib, ptr

BOXDIFF_ATTN_KEYS = [("down", 2, 0, 0), ("down", 2, 1, 0), ("up", 1, 0, 0), ("up", 1, 1, 0), ("up", 1, 2, 0)]
TERM_DTYPE = np.dtype([("tok", "<i4"), ("mask", "<i4"), ("k_fg", "<i4"), ("k_bg", "<i4"), ("corner", "<i4")])


class BoxdiffC(ctypes.Structure):
    """b200lmd_boxdiff (include/b200lmd.h)"""
    _fields_ = [("maps", ctypes.c_void_p * 8), ("dp_extra", ctypes.c_void_p * 8), ("n_keys", ctypes.c_int),
                ("heads", ctypes.c_int), ("n", ctypes.c_int), ("side", ctypes.c_int), ("T", ctypes.c_int),
                ("ext_ld", ctypes.c_int), ("img_term_off", ctypes.c_void_p), ("terms", ctypes.c_void_p),
                ("masks", ctypes.c_void_p), ("corner", ctypes.c_void_p), ("mean", ctypes.c_void_p),
                ("dA", ctypes.c_void_p), ("loss", ctypes.c_void_p), ("kern", ctypes.c_float * 9),
                ("smooth", ctypes.c_int), ("out_scale", ctypes.c_float)]


def gaussian_kernel(kernel_size=3, sigma=0.5):
    """utils/attn.py:89-115 GaussianSmoothing weights: product of exp(-((x - mean) / (2 sigma))^2) per axis, sum 1
    (float32 arithmetic like the reference's torch ops)"""
    ax = np.arange(kernel_size, dtype=np.float32)
    mean = np.float32((kernel_size - 1) / 2)
    g = (np.float32(1.0) / np.float32(sigma * math.sqrt(2 * math.pi))) * \
        np.exp(-(((ax - mean) / np.float32(2 * sigma)) ** 2)).astype(np.float32)
    k = (g[:, None] * g[None, :]).astype(np.float32)
    return (k / k.sum()).astype(np.float32)


@dataclass
class BoxDiffSpec:
    """semantic_guidance_kwargs of generation/boxdiff.py:112-124 + the defaults of utils/boxdiff.py:164,190"""
    layouts: List[G.SampleLayout]
    keys: list = field(default_factory=lambda: list(BOXDIFF_ATTN_KEYS))
    max_index_step: int = 25
    P: float = 0.2
    L: int = 1
    smooth_attentions: bool = True
    sigma: float = 0.5
    kernel_size: int = 3
    amp_loss_scale: float = 10.0
    latent_scale: float = 20.0
    scale_range: tuple = (1.0, 0.5)


def build_tables(layouts, side, P, L):
    """numpy tables (term_off, terms, masks, corner) of a batch at map resolution side x side"""
    n = side * side
    masks, corners, terms, off = [], [], [], [0]
    for b, lay in enumerate(layouts):
        for o, obj_boxes in enumerate(lay.bboxes):
            boxes = G._box_list(obj_boxes)
            m = np.zeros((side, side), dtype=np.uint8)
            cx, cy = np.zeros(side, dtype=np.uint8), np.zeros(side, dtype=np.uint8)
            for box in boxes:
                x0, y0, x1, y1 = G.scale_proportion(box, side, side)
                m[y0:y1, x0:x1] = 1
                cx[max(x0 - L, 0):min(x0 + L + 1, side)] = 1
                cx[max(x1 - L, 0):min(x1 + L + 1, side)] = 1
                cy[max(y0 - L, 0):min(y0 + L + 1, side)] = 1
                cy[max(y1 - L, 0):min(y1 + L + 1, side)] = 1
            s = np.float32(m.sum())
            k_fg = int(np.float32(s * np.float32(P)))               # (obj_mask.sum() * P).long(), boxdiff.py:82
            k_bg = int(np.float32(np.float32(n - s) * np.float32(P)))
            if k_fg < 1 or k_bg < 1:
                raise ValueError(f"BoxDiff: image {b} phrase {o}: box covers {int(s)} of {n} cells, top-k size is 0 "
                                 "(the reference takes the mean of an empty top-k here and produces NaN)")
            mid, cid = len(masks), len(corners)
            masks.append(m.reshape(-1))
            corners.append(np.concatenate([cx, cy]))
            for tok in lay.object_positions[o]:
                terms.append((tok, mid, k_fg, k_bg, cid))
        off.append(len(terms))
    terms_np = np.array(terms, dtype=TERM_DTYPE) if terms else np.zeros(0, dtype=TERM_DTYPE)
    masks_np = np.stack(masks) if masks else np.zeros((1, n), dtype=np.uint8)
    corner_np = np.stack(corners) if corners else np.zeros((1, 2 * side), dtype=np.uint8)
    return np.array(off, dtype=np.int32), terms_np, masks_np, corner_np


class _Holder:
    """per-key dP_extra buffer handed to the backward kernel (no in-kernel loss: c is None)"""
    c = None

    def __init__(self, BH, n, dev, ext_ld):
        self.dp_extra = torch.zeros(BH, n, ext_ld, device=dev, dtype=torch.float32)


class BoxDiffLoss:
    """device tables + scratch of a batch; `holders` goes to B200UNet.guidance_forward as its `losses` argument"""

    def __init__(self, net, spec: BoxDiffSpec, H, W, T, ext_ld=80):
        from .pipelines import _heads_of, _tokens_of
        dev = net.dev
        self.net, self.spec, self.B, self.T, self.ext_ld = net, spec, len(spec.layouts), T, ext_ld
        ns = {_tokens_of(net, k, H, W) for k in spec.keys}
        hs = {_heads_of(net, k) for k in spec.keys}
        if len(ns) != 1 or len(hs) != 1:
            raise ValueError("BoxDiff: all guidance keys must share one map resolution and head count "
                             "(torch.cat over keys, utils/boxdiff.py:146)")
        self.n, self.heads = ns.pop(), hs.pop()
        self.side = int(round(self.n ** 0.5))
        # fixed-capacity device tables refilled in place by update(): their addresses live in captured CUDA graphs
        self.cap_terms, self.cap_masks = self.B * 64, self.B * 16
        self.term_off = torch.zeros(self.B + 1, dtype=torch.int32, device=dev)
        self.terms = torch.zeros(self.cap_terms * TERM_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        self.masks = torch.zeros(self.cap_masks, self.n, dtype=torch.uint8, device=dev)
        self.corner = torch.zeros(self.cap_masks, 2 * self.side, dtype=torch.uint8, device=dev)
        self.holders = {k: _Holder(self.B * self.heads, self.n, dev, ext_ld) for k in spec.keys}
        self.mean = torch.empty(self.B, self.n, T, device=dev, dtype=torch.float32)
        self.dA = torch.empty(self.B, self.n, T, device=dev, dtype=torch.float32)
        self.loss = torch.zeros(self.B, device=dev, dtype=torch.float32)
        if spec.smooth_attentions and spec.kernel_size != 3:
            raise NotImplementedError("BoxDiff smoothing: 3x3 kernel only (the reference's default)")
        if not self.update(spec):
            raise ValueError("BoxDiff: more than 64 phrase tokens or 16 phrases per image")

    def update(self, spec: BoxDiffSpec):
        """refill the tables for a new batch of layouts of the same shape; False when they do not fit the capacity or
        the structure (batch size, keys, smoothing) differs - the caller then builds a fresh object and re-captures"""
        if len(spec.layouts) != self.B or list(spec.keys) != list(self.spec.keys) or \
                (spec.smooth_attentions, spec.kernel_size, spec.sigma) != \
                (self.spec.smooth_attentions, self.spec.kernel_size, self.spec.sigma):      # travel by value in the launch
            return False
        for lay in spec.layouts:
            for pos in lay.object_positions:
                for tok in pos:
                    if not 1 <= tok <= self.T - 2:
                        raise ValueError(f"BoxDiff: token index {tok} outside 1..{self.T - 2} (first and last token are dropped)")
        off, terms, masks, corner = build_tables(spec.layouts, self.side, spec.P, spec.L)
        if len(terms) > self.cap_terms or len(masks) > self.cap_masks:
            return False
        self.spec = spec
        self.term_off.copy_(torch.from_numpy(off))
        if len(terms):
            raw = torch.from_numpy(np.ascontiguousarray(terms).view(np.uint8).reshape(-1).copy())
            self.terms[:raw.numel()].copy_(raw)
        self.masks[:masks.shape[0]].copy_(torch.from_numpy(masks))
        self.corner[:corner.shape[0]].copy_(torch.from_numpy(corner))
        self.kern = gaussian_kernel(spec.kernel_size, spec.sigma)
        return True

    def launch(self, saved):
        """saved: the dict B200UNet filled during guidance_forward: key -> {"probs": fp16 [B, heads, n, T]}"""
        c = BoxdiffC()
        for i, k in enumerate(self.spec.keys):
            c.maps[i] = saved[k]["probs"].data_ptr()
            c.dp_extra[i] = self.holders[k].dp_extra.data_ptr()
        c.n_keys, c.heads, c.n, c.side, c.T, c.ext_ld = len(self.spec.keys), self.heads, self.n, self.side, self.T, self.ext_ld
        c.img_term_off, c.terms = self.term_off.data_ptr(), self.terms.data_ptr()
        c.masks, c.corner = self.masks.data_ptr(), self.corner.data_ptr()
        c.mean, c.dA, c.loss = self.mean.data_ptr(), self.dA.data_ptr(), self.loss.data_ptr()
        for i in range(9):
            c.kern[i] = float(self.kern.reshape(-1)[i])
        c.smooth = int(self.spec.smooth_attentions)
        c.out_scale = self.net.gscale / (len(self.spec.keys) * self.heads)
        check(lib().b200lmd_boxdiff_loss(ctypes.byref(c), ctypes.c_int(self.B), cur_stream()))
        self._keep = (c, saved)

    def gradient_launch(self, z, t_dev, kv_cond, objs=None, fuser_on=False):
        """launch-only (CUDA-graph capturable): returns (gscale * d loss / dz as fp32 NHWC-8 [B, HW, 8], loss [B])"""
        saved = {}
        save = dict(keys=list(self.spec.keys), probs=True, tok=None, out=saved)
        tape, _ = self.net.guidance_forward(z, t_dev, kv_cond, self.holders, objs=objs, fuser_on=fuser_on, save=save)
        self.launch(saved)
        return self.net.guidance_backward(tape), self.loss


def step_scale(spec: BoxDiffSpec, index, n_timesteps):
    """utils/boxdiff.py:228-232: latents -= latent_scale * sqrt(lerp(scale_range, index/(len-1))) / amp * d(loss*amp)/dz
    = latent_scale * sqrt(...) * d loss / dz"""
    s0, s1 = spec.scale_range
    return spec.latent_scale * (s0 + (s1 - s0) * index / (n_timesteps - 1)) ** 0.5
