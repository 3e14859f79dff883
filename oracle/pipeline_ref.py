"""ORACLE (test infrastructure, never on the product path): CPU fp32 restatement of the denoising loops.

Restates models/pipelines.py:16-82 (latent_backward_guidance), :129-247 (generate_semantic_guidance), :324-473
(generate_gligen), :541-599 (generate_partial_frozen) and the DDIM scheduler arithmetic of diffusers 0.18.0
schedulers/scheduling_ddim.py (eta = 0).  One `denoise` routine covers the three loops; behaviours that parity depends
on are kept and cited inline.
"""
from dataclasses import dataclass, field
from typing import Callable, List, Optional

import numpy as np
import torch

from . import guidance_ref, unet_ref


class DDIM:
    """scaled_linear betas 0.00085..0.012, 1000 train steps, steps_offset 1, set_alpha_to_one False."""

    def __init__(self, prediction_type="epsilon"):
        betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = self.alphas_cumprod[0]
        self.prediction_type = prediction_type
        self.num_inference_steps = None
        self.timesteps = None

    def set_timesteps(self, n):
        self.num_inference_steps = n
        ratio = 1000 // n
        self.timesteps = (np.arange(0, n) * ratio).round()[::-1].astype(np.int64) + 1

    def apply_fast_schedule(self, fast_after_steps, fast_rate=2):
        """utils/schedule.py:4-9"""
        ts = self.timesteps
        if fast_after_steps >= len(ts) - 1:
            return
        self.timesteps = np.concatenate([ts[:fast_after_steps], ts[fast_after_steps + 1::fast_rate]])

    def adjust(self, index, t):
        """utils/schedule.py:11-13 (the warnings of :14-19 do not change results)"""
        prev_t = int(self.timesteps[index + 1]) if index + 1 < len(self.timesteps) else -1
        self.num_inference_steps = 1000 // (int(t) - prev_t)

    def step(self, eps_or_v, t, x):
        prev_t = int(t) - 1000 // self.num_inference_steps
        a_t = self.alphas_cumprod[int(t)]
        a_p = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        if self.prediction_type == "epsilon":
            x0 = (x - (1 - a_t) ** 0.5 * eps_or_v) / a_t ** 0.5
            eps = eps_or_v
        else:  # v_prediction
            x0 = a_t ** 0.5 * x - (1 - a_t) ** 0.5 * eps_or_v
            eps = a_t ** 0.5 * eps_or_v + (1 - a_t) ** 0.5 * x
        return a_p ** 0.5 * x0 + (1 - a_p) ** 0.5 * eps


@dataclass
class GuidanceCfg:
    """semantic_guidance_kwargs of the reference (generation/lmd_plus.py:300-320, backward_guidance.py:43-60)"""
    bboxes: list
    object_positions: list
    keys: list
    loss_scale: float = 30.0
    loss_threshold: float = 0.2
    max_iter: object = 5           # int or per-step list
    max_index_step: int = 10
    fg_top_p: float = 0.2
    bg_top_p: float = 0.2
    fg_weight: float = 1.0
    bg_weight: float = 1.0
    ref_maps: Optional[list] = None        # ref_maps[obj][box][step][key] -> [heads, n]
    word_token_indices: Optional[list] = None
    ref_ca_loss_weight: float = 1.0
    ref_word_token_only: bool = False
    use_ratio_based_loss: bool = False     # utils/guidance.py:122-128, the default of compute_ca_lossv3's **kwargs


def guidance_iterations(unet: Callable, sched: DDIM, z, t, index, loss, g: GuidanceCfg, trace=None):
    """models/pipelines.py:16-82.  `loss` carries over between steps (stale-loss loop entry, pipelines.py:30): once
    loss/scale <= threshold no later step ever recomputes it.  `unet(z, t, save_keys)` -> dict key -> [1,heads,n,T]."""
    it = 0
    if index < g.max_index_step:
        mi = g.max_iter
        if isinstance(mi, list):
            mi = mi[index] if len(mi) > index else mi[-1]
        while loss / g.loss_scale > g.loss_threshold and it < mi:
            z = z.detach().requires_grad_(True)
            saved = unet(z, t, g.keys)
            one = {k: v[0] for k, v in saved.items()}
            refs = None
            if g.ref_maps is not None:
                refs = [[box[index] for box in obj] for obj in g.ref_maps]
            L = guidance_ref.ca_loss(one, g.bboxes, g.object_positions, g.keys, g.fg_top_p, g.bg_top_p, g.fg_weight,
                                     g.bg_weight, refs, g.word_token_indices, g.ref_ca_loss_weight,
                                     g.ref_word_token_only, use_ratio_based_loss=g.use_ratio_based_loss) * g.loss_scale
            grad = torch.autograd.grad(L, [z])[0]
            scale = (1 - sched.alphas_cumprod[int(t)]) ** 0.5          # pipelines.py:62-69 (DDIM has no sigmas)
            z = (z - scale * grad).detach()
            loss = float(L)
            it += 1
            if trace is not None:
                trace.append((index, it, loss))
    return z, loss, it


def denoise(w, cfg: unet_ref.UNetConfig, z0, uncond, cond, steps, guidance_scale=7.5, g: Optional[GuidanceCfg] = None,
            frozen_mask=None, frozen_latents=None, frozen_steps=0, gligen=None, gligen_beta=0.3,
            save_keys=None, save_token=None, prediction_type="epsilon", trace=None, fast_after_steps=None,
            fast_rate=2, dynamic_num_inference_steps=False, boxdiff=None):
    """z0 [1,4,H,W]; uncond/cond [1,T,ctx].  boxdiff = dict(bboxes, object_positions, keys, max_index_step[, P, L,
    smooth_attentions]) switches the guidance to utils/boxdiff.py:190-259 (one step per denoising step, sqrt schedule).  Returns dict(latents, latents_all [steps+1], saved (per step), iters).
    gligen: dict(boxes [1,30,4], masks [1,30], positive_embeddings [1,30,768]) for the conditional half.
    Reference quirks reproduced: CFG batch order [uncond; cond] (pipelines.py:420, models.py:85); the guidance pass
    is cond-only and, in GLIGEN mode, sees the ZEROED grounding mask (pipelines.py:317,382-384); fuser is on for
    index < int(beta*steps) (pipelines.py:408-414); frozen blend uses latents_all_input[index+1] (pipelines.py:446)."""
    sched = DDIM(prediction_type)
    sched.set_timesteps(steps)
    if fast_after_steps is not None:                    # pipelines.py:151-152, 358-359
        sched.apply_fast_schedule(fast_after_steps, fast_rate)
    z = z0.clone()
    latents_all = [z.clone()]
    saved_all = []
    iters = []
    loss = 10000.0
    text = torch.cat([uncond, cond], dim=0)
    n_ground = int(gligen_beta * len(sched.timesteps))  # pipelines.py:408
    gl_main = gl_guid = None
    if gligen is not None:
        rep = lambda x: torch.cat([x, x], dim=0)
        masks2 = rep(gligen["masks"]).clone()
        masks2[:1] = 0
        gl_main = dict(boxes=rep(gligen["boxes"]), positive_embeddings=rep(gligen["positive_embeddings"]), masks=masks2)
        gl_guid = dict(boxes=gl_main["boxes"][:1], positive_embeddings=gl_main["positive_embeddings"][:1],
                       masks=gl_main["masks"][:1])
    bd_losses = []
    for index, t in enumerate(sched.timesteps):
        fuser_on = gligen is not None and index < n_ground
        if boxdiff is not None and boxdiff["bboxes"] and index < boxdiff.get("max_index_step", 25):
            from . import boxdiff_ref
            zz = z.detach().requires_grad_(True)
            saved = {}
            unet_ref.unet_forward(w, cfg, zz, t, cond, gligen=gl_guid, fuser_on=fuser_on, saved=saved,
                                  save_keys=boxdiff["keys"])
            L = boxdiff_ref.boxdiff_loss({k: v[0] for k, v in saved.items()}, boxdiff["bboxes"],
                                         boxdiff["object_positions"], boxdiff["keys"], P=boxdiff.get("P", 0.2),
                                         L=boxdiff.get("L", 1),
                                         smooth_attentions=boxdiff.get("smooth_attentions", True)) * 10.0
            grad = torch.autograd.grad(L, [zz])[0]
            z = boxdiff_ref.boxdiff_update(z, grad, index, len(sched.timesteps)).detach()
            bd_losses.append(float(L) / 10.0)
        if g is not None and g.bboxes:
            def guided_unet(zz, tt, keys):
                saved = {}
                unet_ref.unet_forward(w, cfg, zz, tt, cond, gligen=gl_guid, fuser_on=fuser_on, saved=saved,
                                      save_keys=keys)
                return saved
            z, loss, it = guidance_iterations(guided_unet, sched, z, t, index, loss, g, trace)
            iters.append(it)
        with torch.no_grad():
            saved = {} if save_keys is not None else None
            eps = unet_ref.unet_forward(w, cfg, torch.cat([z, z], dim=0), t, text, gligen=gl_main, fuser_on=fuser_on,
                                        saved=saved, save_keys=save_keys)
            if saved is not None:
                # cond half only, single token column (attention_processor.py:466-476)
                saved_all.append({k: (v[1:, :, :, save_token:save_token + 1] if save_token is not None else v[1:])
                                  for k, v in saved.items()})
            eps = eps[:1] + guidance_scale * (eps[1:] - eps[:1])
            if dynamic_num_inference_steps:             # pipelines.py:217-218, 439-440
                sched.adjust(index, t)
            z = sched.step(eps, t, z)
            if frozen_mask is not None and index < frozen_steps:
                z = frozen_latents[index + 1] * frozen_mask + z * (1.0 - frozen_mask)
        if fast_after_steps is None or index < fast_after_steps:   # pipelines.py:449
            latents_all.append(z.clone())
    return dict(latents=z, latents_all=torch.stack(latents_all, 0), saved=saved_all, iters=iters, loss=loss,
                boxdiff_losses=bd_losses)
