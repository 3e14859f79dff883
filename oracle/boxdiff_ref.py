"""ORACLE (test infrastructure, never on the product path): CPU fp32 restatement of the reference's BoxDiff loss
 - groundwork for the BoxDiff kernel, which is not built yet.

Restates utils/boxdiff.py:20-101 (_compute_max_attention_per_index), :104-117 (_compute_loss), :120-161
(compute_ca_loss_boxdiff, without the optional reference-attention term), :164-187
(add_ca_loss_per_attn_map_to_loss_boxdiff) and the update rule of :190-259 (latent_backward_guidance_boxdiff), plus
utils/attn.py:73-131 (GaussianSmoothing).  Pinned against the unmodified reference in
tests/test_oracle_vs_reference.py::test_boxdiff_loss_and_grad_match_reference.
"""
import math

import torch
import torch.nn.functional as F

from . import guidance_ref


BOXDIFF_KEYS = [("down", 2, 0, 0), ("down", 2, 1, 0), ("up", 1, 0, 0), ("up", 1, 1, 0), ("up", 1, 2, 0)]   # the 16x16 maps


def boxdiff_inputs(seed):
    """seeded synthetic case shared by oracle/make_goldens.py and tests/test_oracle_golden.py"""
    g = torch.Generator().manual_seed(100 + seed)
    maps = {k: torch.softmax(2 * torch.randn(8, 256, 77, generator=g), dim=-1) for k in BOXDIFF_KEYS}
    bboxes = [[(0.1, 0.2, 0.6, 0.7)], [(0.5, 0.4, 0.95, 0.9), (0.0, 0.0, 0.3, 0.3)]]
    return maps, bboxes, [[2, 3], [6]]


def gaussian_kernel(kernel_size=3, sigma=0.5):
    """utils/attn.py:89-115: product of per-axis 'gaussians' exp(-((x - mean) / (2 sigma))^2), normalised to sum 1
    (the exponent is the reference's, not the textbook one)."""
    ax = torch.arange(kernel_size, dtype=torch.float32)
    mean = (kernel_size - 1) / 2
    g = 1.0 / (sigma * math.sqrt(2 * math.pi)) * torch.exp(-((ax - mean) / (2 * sigma)) ** 2)
    k = g[:, None] * g[None, :]
    return k / k.sum()


def boxdiff_loss(saved, bboxes, object_positions, keys, P=0.2, L=1, smooth_attentions=True, sigma=0.5, kernel_size=3):
    """saved[key]: [heads, n, T] cross-attention maps (cond half) of the guidance keys (all at the same resolution).
    Average over keys and heads, drop the first and last token, softmax(100 x) over the remaining tokens, then per
    phrase token: inner-box / outer-box top-k means and the corner (projection) L1 terms."""
    maps = torch.cat([saved[k] for k in keys], dim=0).mean(dim=0)          # boxdiff.py:146 [n, T]
    n, T = maps.shape
    H = W = int(math.sqrt(n))
    text = F.softmax(maps.view(H, W, T)[:, :, 1:-1] * 100, dim=-1)          # boxdiff.py:34-36
    kern = gaussian_kernel(kernel_size, sigma)[None, None]
    fg, bg, dx, dy = [], [], [], []
    for obj_idx, positions in enumerate(object_positions):
        obj_boxes = bboxes[obj_idx]
        if not isinstance(obj_boxes[0], (list, tuple)):
            obj_boxes = [obj_boxes]
        for pos in positions:
            image = text[:, :, pos - 1]                                     # token 0 was dropped
            obj_mask = torch.zeros(H, W)
            cmx, cmy = torch.zeros(W), torch.zeros(H)
            for box in obj_boxes:
                x0, y0, x1, y1 = guidance_ref.scale_proportion(box, H, W)
                obj_mask[y0:y1, x0:x1] = 1
                cmx[max(x0 - L, 0):min(x0 + L + 1, W)] = 1.0
                cmx[max(x1 - L, 0):min(x1 + L + 1, W)] = 1.0
                cmy[max(y0 - L, 0):min(y0 + L + 1, H)] = 1.0
                cmy[max(y1 - L, 0):min(y1 + L + 1, H)] = 1.0
            bg_mask = 1 - obj_mask
            if smooth_attentions:
                image = F.conv2d(F.pad(image[None, None], (1, 1, 1, 1), mode="reflect"), kern)[0, 0]
            k = int((obj_mask.sum() * P).long())                            # no max(1, .): k = 0 gives NaN, as in
            fg.append((image * obj_mask).reshape(-1).topk(k)[0].mean())     # the reference
            k = int((bg_mask.sum() * P).long())
            bg.append((image * bg_mask).reshape(-1).topk(k)[0].mean())
            px, py = obj_mask.max(dim=0).values, obj_mask.max(dim=1).values
            dx.append(((image.max(dim=0)[0] - px).abs() * cmx).mean())
            dy.append(((image.max(dim=1)[0] - py).abs() * cmy).mean())
    zero = torch.zeros(())
    loss = sum(torch.maximum(zero, 1.0 - v) for v in fg) + sum(torch.maximum(zero, v) for v in bg) + sum(dx) + sum(dy)
    return loss


def boxdiff_update(z, grad, index, n_timesteps, amp_loss_scale=10.0, latent_scale=20.0, scale_range=(1.0, 0.5)):
    """boxdiff.py:228-232: z <- z - latent_scale * sqrt(lerp(scale_range, index / (len - 1))) / amp_loss_scale * grad,
    grad being d(loss * amp_loss_scale)/dz"""
    scale = (scale_range[0] + (scale_range[1] - scale_range[0]) * index / (n_timesteps - 1)) ** 0.5
    return z - latent_scale * scale / amp_loss_scale * grad
