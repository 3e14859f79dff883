"""BoxDiff baseline - generation/boxdiff.py of the reference: one overall generation with the BoxDiff loss
(inner/outer-box top-k and corner constraints on the layer/head-averaged, re-softmaxed 16x16 maps) for the first
`overall_max_index_step` steps, one guidance step per denoising step.  Keyword surface of generation/boxdiff.py:46;
`run_batch` is the B200 addition (BASELINE config 4 runs 8 images per GPU in lock-step)."""
import torch

from .. import boxdiff as BD
from .. import latents as L
from .. import pipelines as P
from ..guidance import SampleLayout
from . import common
from .common import DEFAULT_OVERALL_NEGATIVE_PROMPT

version = "boxdiff"
height = width = 512
num_inference_steps = 50
guidance_scale = 7.5
overall_guidance_attn_keys = list(BD.BOXDIFF_ATTN_KEYS)       # generation/boxdiff.py:31-37


def run_batch(specs, bg_seeds, overall_max_index_step=25, num_inference_steps=num_inference_steps, height=height,
              width=width, return_latents=False):
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
    spec_bd = BD.BoxDiffSpec(layouts=layouts, keys=list(overall_guidance_attn_keys), max_index_step=overall_max_index_step)
    res = P.denoise(net, torch.cat(z0, 0), torch.cat(uncs, 0), torch.cat(conds, 0), num_inference_steps,
                    guidance_scale=guidance_scale, boxdiff=spec_bd)
    images = env.decode(res["latents"])
    outs = []
    for b in range(len(specs)):
        o = common.Output(image=images[b] if images is not None else None)
        if return_latents:
            o["latents"] = res["latents"][b:b + 1]
            o["guidance_state"] = res["state"]
        outs.append(o)
    return outs


def run(spec, bg_seed=1, overall_max_index_step=25, **kwargs):
    return run_batch([spec], [bg_seed], overall_max_index_step=overall_max_index_step, **kwargs)[0]
