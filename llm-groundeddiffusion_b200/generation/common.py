"""Shared machinery of the generation.* plug-ins: spec conversion, the two-phase layout-grounded generation (per-box
generation -> mask -> latent composition -> overall generation) of generation/lmd.py:215-551 and
generation/lmd_plus.py:193-520, batched over B independent (prompt, layout) pairs.

The B200 UNet, text/VAE/SAM environment are module state set by the driver before use (the reference binds
`models.model_dict` at import, generation/lmd_plus.py:12-19):
    import lgd_b200.generation.common as common
    common.configure(unet=B200UNet(...), env=ReferenceEnv(model_dict))   # or SyntheticEnv()
"""
import os
import sys
import time

import numpy as np
import torch

import ctypes

from .. import latents as L
from .. import pipelines as P
from .._lib import check, cur_stream, lib, ptr
from ..guidance import SampleLayout

DEFAULT_SO_NEGATIVE_PROMPT = ("artifacts, blurry, smooth texture, bad quality, distortions, unrealistic, distorted image, "
                              "bad proportions, duplicate, two, many, group, occlusion, occluded, side, border, collate")
DEFAULT_OVERALL_NEGATIVE_PROMPT = ("artifacts, blurry, smooth texture, bad quality, distortions, unrealistic, "
                                   "distorted image, bad proportions, duplicate")

_state = {"unet": None, "env": None}


class Output(dict):
    """EasyDict-style result (attribute + key access), generate.py reads `.image`"""
    __getattr__ = dict.get


def configure(unet, env):
    _state["unet"], _state["env"] = unet, env


def _need():
    if _state["unet"] is None or _state["env"] is None:
        raise RuntimeError("lgd_b200.generation: call generation.common.configure(unet=..., env=...) first")
    return _state["unet"], _state["env"]


_ONES = ["zero", "one", "two", "three", "four", "five", "six", "seven", "eight", "nine", "ten", "eleven", "twelve",
         "thirteen", "fourteen", "fifteen", "sixteen", "seventeen", "eighteen", "nineteen"]
_TENS = ["", "", "twenty", "thirty", "forty", "fifty", "sixty", "seventy", "eighty", "ninety"]


def _number_words(n):
    """inflect.engine().number_to_words for the counts a layout can hold (utils/parse.py:342)"""
    try:
        import inflect
        return inflect.engine().number_to_words(n)
    except Exception:
        if n < 20:
            return _ONES[n]
        if n < 100:
            return _TENS[n // 10] + ("-" + _ONES[n % 10] if n % 10 else "")
        return str(n)


def _plural(noun):
    try:
        import inflect
        return inflect.engine().plural_noun(noun)
    except Exception:
        if noun.endswith(("s", "x", "z", "ch", "sh")):
            return noun + "es"
        if noun.endswith("y") and noun[-2:-1] not in "aeiou":
            return noun[:-1] + "ies"
        return noun + "s"


def convert_spec(spec, height=512, width=512):
    """utils/parse.py:313-367: boxes sorted by name, xywh(512-space) -> normalised xyxy, per-box prompts
    "<bg> with <name>", repeated objects merged into "<count> <plural>" phrases with several boxes."""
    boxes = sorted(spec["gen_boxes"], key=lambda gb: gb[0])
    conv = []
    for name, b in boxes:
        x0, y0 = b[0] / width, b[1] / height
        conv.append((name, (x0, y0, x0 + b[2] / width, y0 + b[3] / height)))
    bg = spec["bg_prompt"]
    so = [((f"{bg} with {n}" if bg else n), n, n.split(" ")[-1], box) for n, box in conv]
    names = [n for n, _ in conv]
    uniq = sorted(set(names))
    overall = []
    for n in uniq:
        bxs = [box for nn, box in conv if nn == n]
        if len(bxs) > 1:
            ph = _plural(n.replace("an ", "").replace("a ", ""))
            ph = _number_words(len(bxs)) + " " + ph
        else:
            ph = n
        overall.append((ph, ph.split(" ")[-1], bxs))
    objs = ", ".join(p for p, _, _ in overall)
    prompt = (f"{bg} with {objs}" if bg else objs) if objs else bg
    return so, prompt, overall


def centered_box(box, horizontal_only=True, vertical_placement="centered", floor_padding=None):
    """utils/utils.py:19-44 get_centered_box"""
    x0, y0, x1, y1 = box
    w = x1 - x0
    nx0, nx1 = 0.5 - w / 2, 0.5 + w / 2
    if horizontal_only:
        return [nx0, y0, nx1, y1]
    h = y1 - y0
    if vertical_placement == "centered":
        return [nx0, 0.5 - h / 2, nx1, 0.5 + h / 2]
    ny1 = 1 - floor_padding
    return [nx0, ny1 - h, nx1, ny1]


def _gligen_inputs(env, boxes_per_sample, phrases_per_sample, ctx=768, max_objs=30):
    """models/pipelines.py:285-321 prepare_gligen_condition (conditional half; the CFG duplication happens in denoise)"""
    B = len(boxes_per_sample)
    boxes = torch.zeros(B, max_objs, 4)
    emb = torch.zeros(B, max_objs, ctx)
    masks = torch.zeros(B, max_objs)
    for b, (bx, ph) in enumerate(zip(boxes_per_sample, phrases_per_sample)):
        n = min(len(bx), max_objs)
        if n:
            boxes[b, :n] = torch.tensor(bx[:n], dtype=torch.float32)
            emb[b, :n] = env.phrase_embeddings(list(ph[:n]))
            masks[b, :n] = 1
    return dict(boxes=boxes, positive_embeddings=emb, masks=masks)


def layout_generation(specs, bg_seeds, fg_seed_starts, *, use_gligen, so_guidance, overall_guidance, num_inference_steps,
                      frozen_step_ratio, so_beta, overall_beta, so_center_box, so_horizontal_center_only,
                      so_vertical_placement, so_floor_padding, fg_blending_ratio, align_with_overall_bboxes,
                      horizontal_shift_only, use_ref_ca, ref_ca_loss_weight, so_negative_prompt, overall_negative_prompt,
                      guidance_scale=7.5, height=512, width=512, keys=None, overall_prompt_overrides=None,
                      return_latents=False, use_fast_schedule=False, sam_attn_key=("down", 2, 1, 0),
                      sam_attn_start=None):
    """Two-phase generation for a batch of specs.  so_guidance / overall_guidance: dict(loss_scale, loss_threshold,
    max_iter, max_index_step, fg_top_p, bg_top_p, fg_weight, bg_weight) or None.
    sam_attn_start: LMD hands SAM the word-token attention of `sam_attn_key` averaged over the steps from this index
    on and over heads (utils/attn.py:9-38 get_token_attnv2, generation/lmd.py:124-147); None = box-prompted SAM (LMD+,
    generation/lmd_plus.py:123-129)."""
    net, env = _need()
    timing = os.environ.get("B200_TIMING")
    marks = []

    def mark(name):
        if timing:
            torch.cuda.synchronize()
            marks.append((name, time.perf_counter()))

    mark("start")
    keys = list(keys or P.DEFAULT_GUIDANCE_ATTN_KEYS)
    H, W = height // 8, width // 8
    steps = num_inference_steps
    frozen_steps = int(steps * min(max(frozen_step_ratio, 0.0), 1.0))
    # generation/lmd_plus.py:360-367: the per-box generations only need the steps whose latents / attention are
    # transferred; after those the timestep list is thinned (the result only feeds the mask refinement)
    fast_after_steps = None
    if use_fast_schedule:
        ov_mis = overall_guidance["max_index_step"] if overall_guidance is not None else 0
        fast_after_steps = max(frozen_steps, ov_mis) if use_ref_ca else frozen_steps
    B = len(specs)
    conv = [convert_spec(s, height, width) for s in specs]
    so_lists, overall_prompts, overall_pwb = [c[0] for c in conv], [c[1] for c in conv], [c[2] for c in conv]
    for b in range(B):
        if overall_prompt_overrides and overall_prompt_overrides[b] and overall_prompt_overrides[b].strip():
            overall_prompts[b] = overall_prompt_overrides[b].strip()
        if so_center_box:
            so_lists[b] = [(p, ph, w, centered_box(bx, so_horizontal_center_only, so_vertical_placement,
                                                   so_floor_padding)) for p, ph, w, bx in so_lists[b]]
    so_neg = [((s["extra_neg_prompt"] + ", ") if s.get("extra_neg_prompt") else "") + so_negative_prompt for s in specs]
    ov_neg = [((s["extra_neg_prompt"] + ", ") if s.get("extra_neg_prompt") else "") + overall_negative_prompt
              for s in specs]

    # ------------------------------------------------------------------ Phase A: every box of every image, one batch
    owner, z_list, unc_list, cond_list, lay_list, tok_list, so_boxes, so_phrases = [], [], [], [], [], [], [], []
    bg_latents = []
    for b in range(B):
        boxes = [it[3] for it in so_lists[b]]
        inp, bg = L.input_latents_for_boxes(bg_seeds[b], fg_seed_starts[b], boxes, fg_blending_ratio, H, W,
                                            net.cfg.in_channels)
        bg_latents.append(bg)
        if not so_lists[b]:
            continue
        unc, cnd = env.encode_prompts([it[0] for it in so_lists[b]], so_neg[b])
        for i, (prompt, phrase, word, box) in enumerate(so_lists[b]):
            pos, widx, _ = env.phrase_indices(prompt, [phrase], [word], add_suffix=False)
            owner.append(b)
            z_list.append(inp[i])
            unc_list.append(unc)
            cond_list.append(cnd[i:i + 1])
            lay_list.append(SampleLayout([list(box)], pos, widx))
            tok_list.append(widx[0])
            so_boxes.append(box)
            so_phrases.append(phrase)
    latents_all_so, masks_so, saved_so, so_imgs = [], [], [], []
    phaseA = None
    mark("phase A inputs (host)")
    if owner and (use_ref_ca or frozen_steps > 0):
        gspec = None
        if so_guidance is not None and so_guidance["max_index_step"] > 0:
            gspec = P.GuidanceSpec(layouts=lay_list, keys=keys, **so_guidance)
        gl = _gligen_inputs(env, [[b] for b in so_boxes], [[p] for p in so_phrases]) if use_gligen else None
        resA = P.denoise(net, torch.cat(z_list, 0), torch.cat(unc_list, 0), torch.cat(cond_list, 0), steps,
                         guidance_scale=guidance_scale, guidance=gspec, gligen=gl, gligen_beta=so_beta,
                         save_keys=[("down", 2, 1, 0)] + (keys if use_ref_ca else []), save_tok=tok_list,
                         save_latents=True, fast_after_steps=fast_after_steps, dynamic_num_inference_steps=True)
        mark("phase A denoise")
        imgs = env.decode(resA["latents"])
        la_dev = resA["latents_all"]                           # [steps+1, BA, C, H, W], stays on the GPU (f-2)
        tok_attn = None
        if sam_attn_start is not None:
            # utils/attn.py:9-38: mean over the saved steps [start:] then over heads, reshaped to the key's grid
            st = torch.stack([s_[sam_attn_key] for s_ in resA["saved"][sam_attn_start:]], 0).float().mean(0).mean(1)
            side = int(round(st.shape[1] ** 0.5))
            tok_attn = st.view(-1, side, side).cpu().numpy()
        phaseA = dict(latents=resA["latents"], state=resA["state"], token_attn=tok_attn)
        for i in range(len(owner)):
            img = imgs[i] if imgs is not None else None
            so_imgs.append(img)
            masks_so.append(torch.as_tensor(env.refine_mask(
                img, so_boxes[i], H, W, token_attn=tok_attn[i] if tok_attn is not None else None)).bool())
            latents_all_so.append(i)                                # index into la_dev's batch dimension
            saved_so.append([{k: st[k][i] for k in st} for st in resA["saved"]])     # per step {key: [heads, n]}

    mark("phase A outputs to host, masks")
    # ------------------------------------------------------------------ composition
    # host: integer bookkeeping only (mask centres -> cell shifts, ownership maps - like the loss tables);
    # device: one gather kernel builds the composed latents of every image and step from the per-box trajectories
    S_out = (steps if fast_after_steps is None else fast_after_steps) + 1
    BA = len(owner)
    shifts = np.zeros((max(BA, 1), 2), dtype=np.int32)
    owner_maps = torch.zeros(B, H, W, dtype=torch.int32)
    bowner_maps = torch.zeros(B, H, W, dtype=torch.int32)
    composed, frozen_masks, ref_maps, layouts, uncs, conds, glb, glp = [], [], [], [], [], [], [], []
    for b in range(B):
        idx = [i for i, o in enumerate(owner) if o == b] if latents_all_so else []
        msk_b = [masks_so[i] for i in idx]
        phrases = [p for p, _, _ in overall_pwb[b]]
        words = [w for _, w, _ in overall_pwb[b]]
        bboxes = [bx for _, _, bx in overall_pwb[b]]
        flat = [bx for group in bboxes for bx in group]
        offsets = [(0.0, 0.0)] * len(idx)
        if align_with_overall_bboxes and idx:
            # utils/latents.py:85-105 align_with_bboxes: mask centre -> box centre, in multiples of the coarsest grid
            offsets, shifted = [], []
            for m, box, gi in zip(msk_b, flat, idx):
                cx, cy = L.mask_center(m)
                dx = (box[0] + box[2]) / 2 - cx
                dy = 0.0 if horizontal_shift_only else (box[1] + box[3]) / 2 - cy
                offsets.append((dx, dy))
                shifted.append(L.shift(m, dx, dy).bool())
                shifts[gi] = L.shift_cells(dx, dy, H, W)
            msk_b = shifted
        if idx:
            ow, bow = L.compose_owners([m.bool() for m in msk_b])
            gidx = torch.tensor([0] + [i + 1 for i in idx], dtype=torch.int32)
            owner_maps[b], bowner_maps[b] = gidx[ow.long()], gidx[bow.long()]
        frozen_masks.append((owner_maps[b] != 0).float())
        pos, widx, prompt = env.phrase_indices(overall_prompts[b], phrases, words, add_suffix=True)
        unc, cnd = env.encode_prompts([prompt], ov_neg[b])
        uncs.append(unc)
        conds.append(cnd)
        layouts.append(SampleLayout([list(map(tuple, g)) for g in bboxes], pos, widx))
        refs_b = None
        if use_ref_ca and idx:
            refs_b, fi = [], 0
            for group in bboxes:
                per_phrase = []
                for _ in group:
                    steps_maps = saved_so[idx[fi]]
                    if align_with_overall_bboxes:
                        dx, dy = offsets[fi]
                        shifted = []
                        for st in steps_maps:
                            d = {}
                            for k in keys:
                                m = st[k]
                                side = int(round(m.shape[1] ** 0.5))
                                d[k] = L.shift(m.view(m.shape[0], side, side), dx, 0.0 if horizontal_shift_only else dy)\
                                    .reshape(m.shape[0], -1)
                            shifted.append(d)
                        steps_maps = shifted
                    per_phrase.append(steps_maps)
                    fi += 1
                refs_b.append(per_phrase)
        ref_maps.append(refs_b)
        glb.append(flat)
        glp.append([p for p, _, g in overall_pwb[b] for _ in g])

    # ------------------------------------------------------------------ Phase B: overall generation, B images
    gspec = None
    if overall_guidance is not None:
        have_refs = use_ref_ca and any(r is not None for r in ref_maps)
        gspec = P.GuidanceSpec(layouts=layouts, keys=keys, ref_ca_loss_weight=ref_ca_loss_weight,
                               ref_word_token_only=True, ref_maps=ref_maps if have_refs else None, **overall_guidance)
    gl = _gligen_inputs(env, glb, glp) if use_gligen else None
    bg_all = torch.cat(bg_latents, 0).to(net.dev, torch.float32).contiguous()          # [B, C, H, W]
    comp_all = torch.empty(S_out, B, bg_all.shape[1], H, W, device=net.dev, dtype=torch.float32)
    if BA:
        # locals keep the uploaded tables alive across the launch (a temporary passed to ptr() is freed - and its
        # block reused by the next upload - before the kernel runs)
        ow_dev, bow_dev = owner_maps.to(net.dev), bowner_maps.to(net.dev)
        sh_dev = torch.from_numpy(shifts).to(net.dev)
        check(lib().b200lmd_compose_latents(
            ptr(la_dev), ptr(bg_all), ptr(ow_dev), ptr(bow_dev), ptr(sh_dev), ptr(comp_all), ctypes.c_int(S_out),
            ctypes.c_int(BA), ctypes.c_int(B), ctypes.c_int(bg_all.shape[1]), ctypes.c_int(H), ctypes.c_int(W),
            cur_stream()))
        torch.cuda.current_stream().synchronize()      # ... and until it has finished
    else:
        comp_all.zero_()
        comp_all[0] = bg_all
    mark("composition")
    resB = P.denoise(net, comp_all[0], torch.cat(uncs, 0), torch.cat(conds, 0), steps, guidance_scale=guidance_scale,
                     guidance=gspec, frozen_mask=torch.stack(frozen_masks, 0), frozen_latents=comp_all,
                     frozen_steps=frozen_steps, gligen=gl, gligen_beta=overall_beta)
    mark("phase B denoise")
    if timing:
        print("[timing] " + ", ".join(f"{n}: {1e3 * (t - marks[i][1]):.0f} ms" for i, (n, t) in enumerate(marks[1:])),
              file=sys.stderr)
    images = env.decode(resB["latents"])
    outs = []
    for b in range(B):
        idx = [i for i, o in enumerate(owner) if o == b] if so_imgs else []
        o = Output(image=images[b] if images is not None else None, so_img_list=[so_imgs[i] for i in idx])
        if return_latents:
            o["latents"] = resB["latents"][b:b + 1]
            o["guidance_state"] = resB["state"]
            o["phase_a"] = dict(index=idx, masks=[masks_so[i] for i in idx], **(phaseA or {}))
        outs.append(o)
    return outs
