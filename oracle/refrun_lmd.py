"""Drive the UNMODIFIED reference plug-ins generation/lmd.py:215-551 and generation/lmd_plus.py:193-520 (`run`) on CPU -
build container only (needs /root/reference).  TEST INFRASTRUCTURE: used by oracle/make_goldens_layout.py to mint the
golden fixtures of the two-phase layout-grounded generation (BASELINE config 1 and its LMD+ counterpart) and by
tests/test_oracle_vs_reference.py.

What is substituted (nothing in the reference files is edited; module attributes are rebound the way generate.py binds
`models.model_dict`, generate.py:119-127):
  * models.model_dict   = seeded-weight reference UNet + the offline fakes of oracle/fakes.py (word tokenizer, seeded
                          text encoder, toy VAE) + the DDIM scheduler stand-in of oracle/shim/diffusers
  * models.sam.sam_refine_attn / sam_refine_box = "SAM" stand-in: the box raster at the latent resolution
                          (utils.proportion_to_mask), confidence 1.0 - SAM weights do not exist offline
  * recorders wrapped around pipelines.generate_* and guidance.compute_ca_lossv3 (pure pass-through) to capture final
    latents, per-generation loss traces / iteration counts, the SAM inputs (token attention, LMD) and the masks.
"""
import importlib
import sys
import types

import numpy as np
import torch

from . import fakes, ref_loader, refrun, unet_ref


class Recorder:
    def __init__(self):
        self.generations = []     # per generate_* call: dict(kind, latents, losses=[...])
        self.sam_inputs = []      # token_attn_np per per-box generation (LMD)
        self.masks = []
        self._cur = None


def run_reference(method, cfg: unet_ref.UNetConfig, w, spec, bg_seed, fg_seed_start, run_kwargs=None,
                  attn_aggregation_step_start=None, num_threads=None):
    """method: 'lmd' | 'lmd_plus'.  Returns (output EasyDict, Recorder)."""
    assert method in ("lmd", "lmd_plus")
    if num_threads:
        torch.set_num_threads(num_threads)
    r = ref_loader.load()
    import models as ref_models               # the reference package
    from easydict import EasyDict
    from diffusers import DDIMScheduler
    fk = fakes.model_dict_fakes(cfg.cross_attention_dim)
    md = EasyDict(vae=fk["vae"], tokenizer=fk["tokenizer"], text_encoder=fk["text_encoder"],
                  unet=refrun.build_reference_unet(cfg, w), scheduler=DDIMScheduler(), dtype=torch.float32)
    ref_models.model_dict = md
    ref_models.models.model_dict = md
    ref_models.models.torch_device = "cpu"
    ref_models.torch_device = "cpu"
    import utils as ref_utils
    import utils.latents as ref_latents
    ref_latents.torch_device = "cpu"
    from models import pipelines, sam
    sam.torch_device = "cpu"
    modname = "generation." + method
    sys.modules.pop(modname, None)
    gen = importlib.import_module(modname)     # binds model_dict at import (generation/lmd.py:12-19)
    if attn_aggregation_step_start is not None and hasattr(gen, "attn_aggregation_step_start"):
        gen.attn_aggregation_step_start = attn_aggregation_step_start

    rec = Recorder()
    H, W = gen.H, gen.W
    # the per-box generations run in the order of the (centred) so boxes: recompute them the way run() does
    so_list, _, _ = ref_utils.parse.convert_spec(spec, gen.height, gen.width)
    kw = dict(run_kwargs or {})
    import inspect
    defaults = {k: v.default for k, v in inspect.signature(gen.run).parameters.items()}
    getk = lambda k: kw.get(k, defaults.get(k))
    so_boxes = [it[-1] for it in so_list]
    if getk("so_center_box"):
        ck = dict(horizontal_center_only=getk("so_horizontal_center_only"))
        if method == "lmd":
            ck.update(vertical_placement=getk("so_vertical_placement"), floor_padding=getk("so_floor_padding"))
        so_boxes = [ref_utils.get_centered_box(b, **ck) for b in so_boxes]
    box_queue = list(so_boxes)

    def fake_refine_attn(sam_input_image, token_attn_np, model_dict, **k):
        rec.sam_inputs.append(np.array(token_attn_np, dtype=np.float32))
        m = ref_utils.proportion_to_mask(box_queue.pop(0), H, W, return_np=True).astype(bool)
        rec.masks.append(m.copy())
        return m, 1.0

    def fake_refine_box(sam_input_image, box, model_dict, **k):
        m = ref_utils.proportion_to_mask(box, H, W, return_np=True).astype(bool)
        rec.masks.append(m.copy())
        return m, 1.0

    orig = dict(attn=sam.sam_refine_attn, box=sam.sam_refine_box, loss=r.guidance.compute_ca_lossv3)
    sam.sam_refine_attn, sam.sam_refine_box = fake_refine_attn, fake_refine_box

    def loss_rec(*a, **k):
        L = orig["loss"](*a, **k)
        if rec._cur is not None:
            rec._cur["losses"].append(float(L))
        return L
    r.guidance.compute_ca_lossv3 = loss_rec

    wrapped = {}
    for name in ("generate_semantic_guidance", "generate_gligen", "generate_partial_frozen"):
        fn = getattr(pipelines, name)
        wrapped[name] = fn

        def make(fn, name):
            def wrapper(*a, **k):
                rec._cur = dict(kind=name, losses=[])
                out = fn(*a, **k)
                rec._cur["latents"] = out[0].detach().clone()
                rec.generations.append(rec._cur)
                rec._cur = None
                return out
            return wrapper
        setattr(pipelines, name, make(fn, name))
    orig_lbg = pipelines.latent_backward_guidance

    def lbg(*a, **k):
        n0 = len(rec._cur["losses"]) if rec._cur is not None else 0
        res = orig_lbg(*a, **k)
        if rec._cur is not None:
            rec._cur.setdefault("iters", []).append(len(rec._cur["losses"]) - n0)
        return res
    pipelines.latent_backward_guidance = lbg
    try:
        out = gen.run(spec=spec, bg_seed=bg_seed, fg_seed_start=fg_seed_start, **kw)
    finally:
        sam.sam_refine_attn, sam.sam_refine_box = orig["attn"], orig["box"]
        r.guidance.compute_ca_lossv3 = orig["loss"]
        pipelines.latent_backward_guidance = orig_lbg
        for name, fn in wrapped.items():
            setattr(pipelines, name, fn)
    return out, rec
