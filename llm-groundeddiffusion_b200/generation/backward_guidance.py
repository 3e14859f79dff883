"""Backward guidance baseline (Chen et al., training-free layout control) - generation/backward_guidance.py of the
reference: one overall generation with attention guidance, no per-box phase, no GLIGEN.  Same keyword surface and
defaults as generation/backward_guidance.py:43-50; `run_batch` is the B200 addition (B specs in lock-step).

Reference behaviours kept:
  * the loss is the RATIO-BASED energy (utils/guidance.py:122-128): the plug-in does not pass use_ratio_based_loss, so
    compute_ca_lossv3's default (True) applies, with fg/bg weights and top-p unused;
  * ref_ca_loss_weight=0.5 is passed but no reference maps exist (ref_ca_saved_attns=None), so that term is absent;
  * the initial latent comes from bg_seed only (utils/latents.py get_scaled_latents), no foreground blending.
"""
import torch

from .. import latents as L
from .. import pipelines as P
from ..guidance import SampleLayout
from . import common
from .common import DEFAULT_OVERALL_NEGATIVE_PROMPT

version = "backward_guidance"
height = width = 512
num_inference_steps = 50
guidance_scale = 7.5


def run_batch(specs, bg_seeds, overall_loss_scale=30, overall_loss_threshold=0.2, overall_max_iter=5,
              overall_max_index_step=10, num_inference_steps=num_inference_steps, height=height, width=width,
              prediction_type="epsilon", return_latents=False):
    net, env = common._need()
    H, W = height // 8, width // 8
    z0, uncs, conds, layouts = [], [], [], []
    for spec, seed in zip(specs, bg_seeds):
        _, prompt, pwb = common.convert_spec(spec, height, width)
        neg = ((spec["extra_neg_prompt"] + ", ") if spec.get("extra_neg_prompt") else "") + DEFAULT_OVERALL_NEGATIVE_PROMPT
        phrases, words, bboxes = [p for p, _, _ in pwb], [w for _, w, _ in pwb], [b for _, _, b in pwb]
        pos, widx, prompt = env.phrase_indices(prompt, phrases, words, add_suffix=True)
        unc, cnd = env.encode_prompts([prompt], neg)
        uncs.append(unc)
        conds.append(cnd)
        z0.append(L.seeded_noise(seed, net.cfg.in_channels, H, W))
        layouts.append(SampleLayout([list(map(tuple, g)) for g in bboxes], pos, widx))
    gspec = P.GuidanceSpec(layouts=layouts, loss_scale=overall_loss_scale, loss_threshold=overall_loss_threshold,
                           max_iter=overall_max_iter, max_index_step=overall_max_index_step, ref_ca_loss_weight=0.5,
                           ref_word_token_only=True, ref_maps=None, use_ratio_based_loss=True)
    res = P.denoise(net, torch.cat(z0, 0), torch.cat(uncs, 0), torch.cat(conds, 0), num_inference_steps,
                    guidance_scale=guidance_scale, guidance=gspec, prediction_type=prediction_type)
    images = env.decode(res["latents"])
    outs = []
    for b in range(len(specs)):
        o = common.Output(image=images[b] if images is not None else None)
        if return_latents:
            o["latents"] = res["latents"][b:b + 1]
            o["guidance_state"] = res["state"]
        outs.append(o)
    return outs


def run(spec, bg_seed=1, overall_loss_scale=30, overall_loss_threshold=0.2, overall_max_iter=5,
        overall_max_index_step=10, **kwargs):
    return run_batch([spec], [bg_seed], overall_loss_scale=overall_loss_scale,
                     overall_loss_threshold=overall_loss_threshold, overall_max_iter=overall_max_iter,
                     overall_max_index_step=overall_max_index_step, **kwargs)[0]
