"""Host mirror of the reference's guidance-loss front end (utils/guidance.py, utils/utils.py:57-70).

The arithmetic of the loss runs inside xattn_fwd_kernel (csrc/xattn.cuh); this module only turns the reference's Python
inputs - bboxes (2- or 3-level lists of normalised xyxy), object_positions (token index lists), word_token_indices,
ref_ca_saved_attns - into the flat integer/float tables the kernel reads.  Every integer here (cell ranges, k_fg, k_bg)
must equal the reference's bit for bit:
  scale_proportion            utils/utils.py:57-70   (Python banker's round on origin and size separately)
  union-of-boxes mask          utils/guidance.py:102-114
  k = max(1, floor(count*p))   utils/guidance.py:136-137 (float32 product, truncation)
  normalisers                  utils/guidance.py:146,237,270,284
"""
import ctypes
import math
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from ._lib import lib

TERM_DTYPE = np.dtype([("type", "<i4"), ("slot", "<i4"), ("mask", "<i4"), ("k_fg", "<i4"), ("k_bg", "<i4"),
                       ("w_fg", "<f4"), ("w_bg", "<f4"), ("w_ref", "<f4"), ("ref", "<i4")])


class XattnLossC(ctypes.Structure):
    """b200lmd_xattn_loss (include/b200lmd.h)"""
    _fields_ = [("img_term_off", ctypes.c_void_p), ("terms", ctypes.c_void_p), ("masks", ctypes.c_void_p),
                ("refs", ctypes.c_void_p), ("slot_tok", ctypes.c_void_p), ("pcol", ctypes.c_void_p),
                ("counters", ctypes.c_void_p), ("loss_part", ctypes.c_void_p), ("dp_extra", ctypes.c_void_p),
                ("ext_ld", ctypes.c_int), ("gscale", ctypes.c_float), ("eps", ctypes.c_float)]


def scale_proportion(box, H, W):
    x_min, y_min = round(box[0] * W), round(box[1] * H)
    bw, bh = round((box[2] - box[0]) * W), round((box[3] - box[1]) * H)
    x_max, y_max = x_min + bw, y_min + bh
    return max(x_min, 0), max(y_min, 0), min(x_max, W), min(y_max, H)


def _box_list(obj_boxes):
    if not isinstance(obj_boxes[0], (list, tuple)):
        return [obj_boxes]
    return list(obj_boxes)


def box_mask(boxes, side):
    m = np.zeros((side, side), dtype=np.uint8)
    for b in boxes:
        x0, y0, x1, y1 = scale_proportion(b, side, side)
        m[y0:y1, x0:x1] = 1
    return m.reshape(-1)


def topk_sizes(mask, fg_top_p, bg_top_p):
    s = np.float32(mask.sum())
    k_fg = max(1, int(np.float32(s * np.float32(fg_top_p))))
    k_bg = max(1, int(np.float32(np.float32(mask.size - s) * np.float32(bg_top_p))))
    return k_fg, k_bg


@dataclass
class SampleLayout:
    """guidance inputs of ONE image, in the reference's own formats"""
    bboxes: list                       # [phrase] -> box or [boxes]
    object_positions: list             # [phrase] -> [token indices]
    word_token_indices: Optional[list] = None
    # [phrase][box] -> {key: array/tensor [heads, n]} for the current step, or a LIST of such dicts over steps
    ref_maps: Optional[list] = None


@dataclass
class LossParams:
    loss_scale: float = 30.0
    fg_top_p: float = 0.2
    bg_top_p: float = 0.2
    fg_weight: float = 1.0
    bg_weight: float = 1.0
    ref_ca_loss_weight: float = 1.0
    ref_word_token_only: bool = False
    use_ref: bool = False
    # utils/guidance.py:122-128: the deprecated ratio-based energy (1 - sum(P M)/sum(P))^2, mean over heads; it is what
    # compute_ca_lossv3 does when the caller does not pass use_ratio_based_loss=False (generation/backward_guidance.py)
    use_ratio_based_loss: bool = False


def assign_slots(samples: Sequence[SampleLayout], params: LossParams):
    """saved-column slots per image: every token any loss term reads, in ascending order"""
    max_slots = lib().b200lmd_max_loss_slots()
    slot_tok = np.full((len(samples), max_slots), -1, dtype=np.int32)
    slot_of = []
    for b, s in enumerate(samples):
        toks = set()
        for o, pos in enumerate(s.object_positions):
            toks.update(pos)
            if params.use_ref and s.ref_maps is not None:
                toks.add(s.word_token_indices[o] if params.ref_word_token_only else pos[-1])
        toks = sorted(toks)
        if len(toks) > max_slots:
            raise ValueError(f"image {b}: {len(toks)} distinct guidance tokens > {max_slots}")
        slot_tok[b, :len(toks)] = toks
        slot_of.append({t: i for i, t in enumerate(toks)})
    return slot_tok, slot_of


def build_key_tables(samples: Sequence[SampleLayout], slot_of, key, n, heads, n_keys, params: LossParams):
    """tables of one attention key (resolution n): returns numpy (term_off, terms, masks, refs)"""
    side = int(math.sqrt(n))
    masks, refs, terms = [], [], []
    term_off = [0]
    S = params.loss_scale
    for b, s in enumerate(samples):
        n_obj = len(s.bboxes)
        for o in range(n_obj):
            boxes = _box_list(s.bboxes[o])
            m = box_mask(boxes, side)
            k_fg, k_bg = topk_sizes(m, params.fg_top_p, params.bg_top_p)
            mid = len(masks)
            masks.append(m)
            T_o = len(s.object_positions[o])
            norm = S / (T_o * n_obj * n_keys)
            for tok in s.object_positions[o]:
                if params.use_ratio_based_loss:
                    terms.append((2, slot_of[b][tok], mid, 1, 1, norm / heads, 0.0, 0.0, 0))
                else:
                    terms.append((0, slot_of[b][tok], mid, k_fg, k_bg, norm * params.fg_weight,
                                  norm * params.bg_weight, 0.0, 0))
        if params.use_ref and s.ref_maps is not None and params.ref_ca_loss_weight != 0.0:
            for o in range(n_obj):
                boxes = _box_list(s.bboxes[o])
                toks = [s.word_token_indices[o]] if params.ref_word_token_only else [s.object_positions[o][-1]]
                w = S * params.ref_ca_loss_weight / (len(boxes) * len(toks)) / (n_obj * n_keys) / heads
                for bi, box in enumerate(boxes):
                    mid = len(masks)
                    masks.append(box_mask([box], side))
                    rid = len(refs)
                    refs.append(s.ref_maps[o][bi])       # resolved per step in KeyLoss.set_step
                    for tok in toks:
                        terms.append((1, slot_of[b][tok], mid, 0, 0, 0.0, 0.0, w, rid))
        if len(terms) - term_off[-1] > 80:
            raise ValueError(f"image {b}: more than 80 loss terms")
        term_off.append(len(terms))
    terms_np = np.array(terms, dtype=TERM_DTYPE) if terms else np.zeros(0, dtype=TERM_DTYPE)
    masks_np = np.stack(masks) if masks else np.zeros((1, n), dtype=np.uint8)
    return np.array(term_off, dtype=np.int32), terms_np, masks_np, refs


class KeyLoss:
    """device-resident loss tables + scratch of one attention key.

    Every device buffer is allocated ONCE at a fixed capacity and refilled in place by `update()`: their addresses are
    baked into captured CUDA graphs (the b200lmd_xattn_loss struct travels by value in the kernel parameters), so a
    graph captured for one batch of layouts is replayed for the next batch after an update - no re-capture."""

    def __init__(self, samples, slot_tok_dev, slot_of, key, n, heads, n_keys, params: LossParams, device, ext_ld=80,
                 gscale=1.0, mask_rows=None, ref_rows=None):
        B = len(samples)
        self.n, self.heads, self.B = n, heads, B
        self.key, self.device = key, device
        max_slots = lib().b200lmd_max_loss_slots()
        off, terms, masks, refs = build_key_tables(samples, slot_of, key, n, heads, n_keys, params)
        self.cap_terms = B * 80
        self.cap_masks = max(len(masks), mask_rows or B * 16)
        self.cap_refs = max(len(refs), ref_rows or B * 8, 1)
        self.term_off = torch.zeros(B + 1, dtype=torch.int32, device=device)
        self.terms = torch.zeros(self.cap_terms * TERM_DTYPE.itemsize, dtype=torch.uint8, device=device)
        self.masks = torch.zeros(self.cap_masks, n, dtype=torch.uint8, device=device)
        # reference maps live in ONE static device buffer; set_step() refills it in place.  Entries may be host arrays
        # or device tensors (Phase-A maps stay on the GPU).
        self.refs = torch.zeros(self.cap_refs, heads, n, device=device, dtype=torch.float32)
        self.slot_tok = slot_tok_dev
        self.pcol = torch.zeros(B * heads, max_slots, n, device=device, dtype=torch.float32)
        self.counters = torch.zeros(B * heads, device=device, dtype=torch.int32)
        self.loss_part = torch.zeros(B * heads, device=device, dtype=torch.float32)
        self.dp_extra = torch.zeros(B * heads, n, ext_ld, device=device, dtype=torch.float32)
        self.c = XattnLossC(self.term_off.data_ptr(), self.terms.data_ptr(), self.masks.data_ptr(),
                            self.refs.data_ptr(), self.slot_tok.data_ptr(), self.pcol.data_ptr(),
                            self.counters.data_ptr(), self.loss_part.data_ptr(), self.dp_extra.data_ptr(), ext_ld,
                            gscale, 1e-5)
        self._fill(off, terms, masks, refs)

    def fits(self, n_terms, n_masks, n_refs):
        return n_terms <= self.cap_terms and n_masks <= self.cap_masks and n_refs <= self.cap_refs

    def _fill(self, off, terms, masks, refs):
        dev = self.device
        self.term_off.copy_(torch.from_numpy(off))
        if len(terms):
            raw = torch.from_numpy(terms.view(np.uint8).reshape(-1).copy())
            self.terms[:raw.numel()].copy_(raw)
        self.masks[:masks.shape[0]].copy_(torch.from_numpy(masks))
        self.ref_entries = refs
        self.set_step(0)

    def update(self, samples, slot_of, n_keys, params: LossParams):
        """refill the static tables for a new batch of layouts (same B, key, resolution); returns False when the new
        tables exceed the allocated capacity (the caller then builds a fresh KeyLoss and re-captures)"""
        if len(samples) != self.B:
            return False
        off, terms, masks, refs = build_key_tables(samples, slot_of, self.key, self.n, self.heads, n_keys, params)
        if not self.fits(len(terms), len(masks), len(refs)):
            return False
        self._fill(off, terms, masks, refs)
        return True

    def set_step(self, index):
        """load the reference maps of denoising step `index` (utils/guidance.py:181) into the static buffer"""
        for i, e in enumerate(self.ref_entries):
            m = (e[index] if isinstance(e, list) else e)[self.key]
            if not torch.is_tensor(m):
                m = torch.from_numpy(np.asarray(m, dtype=np.float32))
            self.refs[i].copy_(m.reshape(self.heads, self.n), non_blocking=True)

    def loss_per_image(self):
        return self.loss_part.view(self.B, self.heads).sum(dim=1)
