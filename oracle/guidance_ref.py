"""ORACLE (test infrastructure, never on the product path): CPU fp32 restatement of the LMD / LMD+ attention loss.

Restates utils/guidance.py:91-286 (compute_ca_lossv3, add_ca_loss_per_attn_map_to_loss max-based branch,
add_ref_ca_loss_per_attn_map_to_lossv2) and utils/utils.py:57-70 (scale_proportion).  Two forms are provided:
  * ca_loss(...)          differentiable torch (autograd supplies d loss / d P for the oracle's guidance step)
  * ca_loss_and_grad(...) numpy closed form of the loss AND d loss / d P, which is what the
                          CUDA backward kernel is compared against, and itself checked against autograd in tests.
Layout convention: P[key] is [heads, n, T] for ONE sample (the reference squeezes batch 1, utils/guidance.py:264).
"""
import math

import numpy as np
import torch


def scale_proportion(box, H, W):
    """utils/utils.py:61-68: origin and size rounded separately with Python's banker's round, then clamped."""
    x_min, y_min = round(box[0] * W), round(box[1] * H)
    bw, bh = round((box[2] - box[0]) * W), round((box[3] - box[1]) * H)
    x_max, y_max = x_min + bw, y_min + bh
    return max(x_min, 0), max(y_min, 0), min(x_max, W), min(y_max, H)


def _as_box_list(obj_boxes):
    """two-level (one box per phrase) vs three-level (several boxes per phrase) input, utils/guidance.py:108-109"""
    if not isinstance(obj_boxes[0], (list, tuple)):
        return [obj_boxes]
    return list(obj_boxes)


def box_mask(boxes, side):
    """union-of-boxes cell mask [side*side] float32 (utils/guidance.py:104-114)"""
    m = np.zeros((side, side), dtype=np.float32)
    for b in boxes:
        x0, y0, x1, y1 = scale_proportion(b, side, side)
        m[y0:y1, x0:x1] = 1.0
    return m.reshape(-1)


def topk_sizes(mask, fg_top_p, bg_top_p):
    """k = max(1, floor(count * p)) computed the way the reference does: float32 product, truncation
    (utils/guidance.py:136-137: (mask.sum() * p).long().clamp_(min=1))."""
    s = np.float32(mask.sum())
    k_fg = max(1, int(np.float32(s * np.float32(fg_top_p))))
    k_bg = max(1, int(np.float32(np.float32(mask.size - s) * np.float32(bg_top_p))))
    return k_fg, k_bg


def ca_loss(saved, bboxes, object_positions, keys, fg_top_p=0.2, bg_top_p=0.2, fg_weight=1.0, bg_weight=1.0,
            ref_maps=None, word_token_indices=None, ref_ca_loss_weight=1.0, ref_word_token_only=False, eps=1e-5,
            use_ratio_based_loss=False):
    """saved[key]: torch [heads, n, T] (one sample).  ref_maps[obj][box][key]: torch [heads, n] at this timestep.
    Returns the scalar torch loss BEFORE multiplication by loss_scale (pipelines.py:48 multiplies afterwards)."""
    n_obj = len(bboxes)
    total = torch.zeros((), dtype=torch.float32)
    if n_obj == 0:
        return total
    for key in keys:
        P = saved[key].float()
        n = P.shape[1]
        side = int(math.sqrt(n))
        for o in range(n_obj):
            m_np = box_mask(_as_box_list(bboxes[o]), side)
            k_fg, k_bg = topk_sizes(m_np, fg_top_p, bg_top_p)
            m = torch.from_numpy(m_np)
            obj = torch.zeros((), dtype=torch.float32)
            for tok in object_positions[o]:
                col = P[:, :, tok]                                             # [heads, n]
                if use_ratio_based_loss:                                       # utils/guidance.py:122-128
                    act = (col * m).sum(dim=-1) / col.sum(dim=-1)
                    obj = obj + torch.mean((1 - act) ** 2)
                    continue
                obj = obj + fg_weight * (1 - (col * m).topk(k_fg, dim=1).values.mean(dim=1)).sum()
                obj = obj + bg_weight * ((col * (1 - m)).topk(k_bg, dim=1).values.mean(dim=1)).sum()
            total = total + obj / len(object_positions[o])
    total = total / (n_obj * len(keys))
    if ref_maps is not None and ref_ca_loss_weight != 0.0:
        ref_total = torch.zeros((), dtype=torch.float32)
        for o in range(n_obj):
            boxes = _as_box_list(bboxes[o])
            toks = [word_token_indices[o]] if ref_word_token_only else [object_positions[o][-1]]
            obj = torch.zeros((), dtype=torch.float32)
            for bi, box in enumerate(boxes):
                for key in keys:
                    P = saved[key].float()
                    side = int(math.sqrt(P.shape[1]))
                    m = torch.from_numpy(box_mask([box], side))
                    R = ref_maps[o][bi][key].float()
                    for tok in toks:
                        a = P[:, :, tok] * m
                        a = a / (a.sum(dim=-1, keepdim=True) + eps)
                        r = R * m
                        r = r / (r.sum(dim=-1, keepdim=True) + eps)
                        obj = obj + (a - r).abs().sum(dim=-1).mean(dim=0)
            ref_total = ref_total + ref_ca_loss_weight * obj / (len(boxes) * len(toks))
        total = total + ref_total / (n_obj * len(keys))
    return total


def ca_loss_and_grad(saved, bboxes, object_positions, keys, fg_top_p=0.2, bg_top_p=0.2, fg_weight=1.0, bg_weight=1.0,
                     ref_maps=None, word_token_indices=None, ref_ca_loss_weight=1.0, ref_word_token_only=False,
                     eps=1e-5):
    """numpy closed form; saved[key]: float32 [heads, n, T].  Returns (loss, {key: dloss/dP [heads, n, T]}).
    Top-k ties are broken towards the lower index (any choice gives the same loss; masked-out cells have zero
    gradient either way because they are multiplied by the mask)."""
    n_obj = len(bboxes)
    grads = {k: np.zeros_like(saved[k], dtype=np.float32) for k in keys}
    if n_obj == 0:
        return 0.0, grads
    loss = 0.0
    norm = 1.0 / (n_obj * len(keys))
    for key in keys:
        P = saved[key].astype(np.float32)
        heads, n, _ = P.shape
        side = int(math.sqrt(n))
        for o in range(n_obj):
            m = box_mask(_as_box_list(bboxes[o]), side)
            k_fg, k_bg = topk_sizes(m, fg_top_p, bg_top_p)
            T_o = len(object_positions[o])
            for tok in object_positions[o]:
                col = P[:, :, tok]
                for msk, k, sign, wgt in ((m, k_fg, -1.0, fg_weight), (1 - m, k_bg, +1.0, bg_weight)):
                    v = col * msk
                    idx = np.argsort(-v, axis=1, kind="stable")[:, :k]
                    top = np.take_along_axis(v, idx, axis=1)
                    mean = top.mean(axis=1)
                    loss += norm / T_o * wgt * ((1 - mean).sum() if sign < 0 else mean.sum())
                    g = np.zeros_like(col)
                    np.put_along_axis(g, idx, 1.0, axis=1)
                    grads[key][:, :, tok] += norm / T_o * wgt * sign / k * g * msk
    if ref_maps is not None and ref_ca_loss_weight != 0.0:
        for o in range(n_obj):
            boxes = _as_box_list(bboxes[o])
            toks = [word_token_indices[o]] if ref_word_token_only else [object_positions[o][-1]]
            wgt = norm * ref_ca_loss_weight / (len(boxes) * len(toks))
            for bi, box in enumerate(boxes):
                for key in keys:
                    P = saved[key].astype(np.float32)
                    heads, n, _ = P.shape
                    m = box_mask([box], int(math.sqrt(n)))
                    R = np.asarray(ref_maps[o][bi][key], dtype=np.float32)
                    for tok in toks:
                        a = P[:, :, tok] * m
                        A = a.sum(axis=1, keepdims=True) + eps
                        ah = a / A
                        r = R * m
                        rh = r / (r.sum(axis=1, keepdims=True) + eps)
                        sgn = np.sign(ah - rh)
                        loss += wgt * np.abs(ah - rh).sum(axis=1).mean()
                        inner = (sgn * ah).sum(axis=1, keepdims=True)
                        grads[key][:, :, tok] += wgt / heads * m / A * (sgn - inner)
    return float(loss), grads
