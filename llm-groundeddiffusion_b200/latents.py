"""Host mirror of utils/latents.py and the geometry helpers of utils/utils.py that feed the denoising path.

Tiny CPU tensor bookkeeping (seeded initial noise, foreground/background blending, box-aligned shifting, per-timestep
masked composition); no arithmetic of the hot path lives here.  Reference behaviours kept (SURVEY.md Appendix B 9,10):
initial noise comes from the CPU generator in float32 (utils/latents.py:13-16), a foreground seed equal to the
background seed is bumped by 12345 (:145-147), composition visits the largest mask first (:58-60).
"""
import numpy as np
import torch

from .guidance import scale_proportion


def seeded_noise(seed, channels, h, w):
    g = torch.Generator().manual_seed(int(seed))
    return torch.randn((1, channels, h, w), generator=g, dtype=torch.float32)


def box_to_mask(box, H, W):
    """utils/utils.py:47-55 proportion_to_mask"""
    x0, y0, x1, y1 = scale_proportion(box, H, W)
    m = torch.zeros(H, W)
    m[y0:y1, x0:x1] = 1.0
    return m


def input_latents_for_boxes(bg_seed, fg_seed_start, boxes, fg_blending_ratio, H, W, channels=4):
    """utils/latents.py:120-160 get_input_latents_list (DDIM: init_noise_sigma = 1)"""
    bg = seeded_noise(bg_seed, channels, H, W)
    out = []
    for i, box in enumerate(boxes):
        fg_seed = fg_seed_start + i
        if fg_seed == bg_seed:
            fg_seed += 12345
        fg = seeded_noise(fg_seed, channels, H, W)
        m = box_to_mask(box, H, W)
        r = fg_blending_ratio
        out.append((bg * (1.0 - m) + (bg * np.sqrt(1.0 - r) + fg * np.sqrt(r)) * m).float())
    return out, bg


def mask_center(mask):
    """utils/utils.py:102-121 binary_mask_to_center(normalize=True)"""
    h, w = mask.shape
    m = mask.float()
    total = m.sum()
    x = float((m.sum(dim=0) @ torch.arange(w).float()) / total)
    y = float((m.sum(dim=1) @ torch.arange(h).float()) / total)
    return x / w, y / h


def mask_to_box_mask(mask):
    """utils/utils.py:72-100: tight box of the mask enlarged by one cell, filled inclusive of its max corner"""
    ys, xs = torch.where(mask)
    h, w = mask.shape
    y0, y1 = max(int(ys.min()) - 1, 0), min(int(ys.max()) + 1, h)
    x0, x1 = max(int(xs.min()) - 1, 0), min(int(xs.max()) + 1, w)
    out = torch.zeros(h, w)
    out[y0:y1 + 1, x0:x1 + 1] = 1.0
    return out


def shift(t, x_off, y_off, base=8, normalized=True, channels_last=False):
    """utils/utils.py:145-180 shift_tensor: integer shift (in multiples of the coarsest grid when normalized), zero fill"""
    hh, ww = (t.shape[-3], t.shape[-2]) if channels_last else (t.shape[-2], t.shape[-1])
    if normalized:
        x_off = round(x_off * base) * (ww // base)
        y_off = round(y_off * base) * (hh // base)
    out = torch.zeros_like(t)
    ow, oh = ww - abs(x_off), hh - abs(y_off)
    ys, yd = (0, y_off) if y_off >= 0 else (-y_off, 0)
    xs, xd = (0, x_off) if x_off >= 0 else (-x_off, 0)
    if channels_last:
        out[..., yd:yd + oh, xd:xd + ow, :] = t[..., ys:ys + oh, xs:xs + ow, :]
    else:
        out[..., yd:yd + oh, xd:xd + ow] = t[..., ys:ys + oh, xs:xs + ow]
    return out


def align_to_boxes(latents_all_list, masks, boxes, horizontal_only=False):
    """utils/latents.py:85-105 align_with_bboxes"""
    out_l, out_m, offs = [], [], []
    for lat, m, box in zip(latents_all_list, masks, boxes):
        cx, cy = mask_center(m)
        dx = (box[0] + box[2]) / 2 - cx
        dy = 0.0 if horizontal_only else (box[1] + box[3]) / 2 - cy
        out_l.append(shift(lat, dx, dy))
        out_m.append(shift(m, dx, dy))
        offs.append((dx, dy))
    return out_l, out_m, offs


def compose(latents_all_list, masks, latents_bg, steps, compose_box_to_bg=True):
    """utils/latents.py:37-83 compose_latents.  latents_all_list[i]: [steps+1, 1, C, H, W]; masks[i]: bool [H, W].
    Returns composed [steps+1, 1, C, H, W] and foreground_indices [H, W] (0 = background)."""
    composed = torch.zeros((steps + 1, *latents_bg.shape), dtype=torch.float32)
    composed[0] = latents_bg
    fg_idx = torch.zeros(latents_bg.shape[-2:], dtype=torch.long)
    order = np.argsort(-np.array([float(m.sum()) for m in masks])) if masks else []
    if compose_box_to_bg:
        for i in order:
            bm = mask_to_box_mask(masks[i])[None, None]
            composed[0] = composed[0] * (1.0 - bm) + latents_all_list[i][0] * bm
    for i in order:
        m = masks[i].bool()
        fg_idx = fg_idx * (~m) + (int(i) + 1) * m
        mf = m[None, None, None].float()
        composed = composed * (1.0 - mf) + latents_all_list[i] * mf
    return composed, fg_idx


def shift_cells(x_off, y_off, hh, ww, base=8):
    """the integer cell shift shift() applies to a normalised offset (utils/utils.py:160-164)"""
    return round(x_off * base) * (ww // base), round(y_off * base) * (hh // base)


def compose_owners(masks):
    """Ownership maps of compose_latents (utils/latents.py:56-78) for already-shifted masks[i] (bool [H, W]):
    owner[y, x] = 1 + i of the last mask in composition order (largest first) covering the cell, bowner the same for the
    enlarged box masks that blend step 0 into the background; 0 = none.  `owner` is the reference's
    foreground_indices."""
    H, W = masks[0].shape if masks else (0, 0)
    owner = torch.zeros((H, W), dtype=torch.int32)
    bowner = torch.zeros((H, W), dtype=torch.int32)
    order = np.argsort(-np.array([float(m.sum()) for m in masks])) if masks else []
    for i in order:
        bm = mask_to_box_mask(masks[i]).bool()
        bowner[bm] = int(i) + 1
    for i in order:
        owner[masks[i].bool()] = int(i) + 1
    return owner, bowner
