"""GLIGEN baseline (adapters only, no attention guidance, no per-box phase) - generation/gligen.py of the reference:
`run(spec, gligen_scheduled_sampling_beta=0.4, bg_seed=1)`; one box + phrase per layout entry (generation/gligen.py:
52-53: the per-box names and boxes, not the merged overall phrases).  `run_batch` is the B200 addition."""
import torch

from .. import latents as L
from .. import pipelines as P
from . import common
from .common import DEFAULT_OVERALL_NEGATIVE_PROMPT

version = "gligen"
height = width = 512
num_inference_steps = 50
guidance_scale = 7.5


def run_batch(specs, bg_seeds, gligen_scheduled_sampling_beta=0.4, num_inference_steps=num_inference_steps,
              height=height, width=width, return_latents=False):
    net, env = common._need()
    if not net.cfg.use_gated_attention:
        raise RuntimeError("generation.gligen needs a GLIGEN UNet (UNetConfig.use_gated_attention)")  # gligen.py:14
    H, W = height // 8, width // 8
    z0, uncs, conds, boxes, phrases = [], [], [], [], []
    for spec, seed in zip(specs, bg_seeds):
        so, prompt, _ = common.convert_spec(spec, height, width)
        neg = ((spec["extra_neg_prompt"] + ", ") if spec.get("extra_neg_prompt") else "") + DEFAULT_OVERALL_NEGATIVE_PROMPT
        unc, cnd = env.encode_prompts([prompt], neg)
        uncs.append(unc)
        conds.append(cnd)
        z0.append(L.seeded_noise(seed, net.cfg.in_channels, H, W))
        boxes.append([list(it[3]) for it in so])
        phrases.append([it[1] for it in so])
    gl = common._gligen_inputs(env, boxes, phrases)
    res = P.denoise(net, torch.cat(z0, 0), torch.cat(uncs, 0), torch.cat(conds, 0), num_inference_steps,
                    guidance_scale=guidance_scale, gligen=gl, gligen_beta=gligen_scheduled_sampling_beta)
    images = env.decode(res["latents"])
    outs = []
    for b in range(len(specs)):
        o = common.Output(image=images[b] if images is not None else None)
        if return_latents:
            o["latents"] = res["latents"][b:b + 1]
        outs.append(o)
    return outs


def run(spec, gligen_scheduled_sampling_beta=0.4, bg_seed=1, **kwargs):
    return run_batch([spec], [bg_seed], gligen_scheduled_sampling_beta=gligen_scheduled_sampling_beta, **kwargs)[0]
