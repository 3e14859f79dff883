"""ORACLE (test infrastructure, never on the product path): the reference's denoising loop restated against the
reference's own UNet CALL SHAPE, so that any object with that operator surface can be driven by it:

    unet(sample, t, encoder_hidden_states=..., cross_attention_kwargs={save_attn_to_dict, save_keys, return_cond_ca_only,
         return_token_ca_only, gligen{boxes, positive_embeddings, masks}}).sample      models/pipelines.py:44,200,427
    unet.modules() -> fuser handles with .enabled                                       models/pipelines.py:280-283

Restates models/pipelines.py:16-82 (latent_backward_guidance: saved maps -> loss -> torch.autograd.grad w.r.t. the
latents) and :324-473 (generate_gligen; without `gligen` it is generate_semantic_guidance :129-247).  Used three ways:
  * build container: driven with the UNMODIFIED reference UNet module and compared with the unmodified
    pipelines.generate_gligen (tests/test_oracle_vs_reference.py) - pins this loop;
  * GPU box: driven with lgd_b200.adapter.B200UNetAdapter and compared with the same loop over `OracleUNet` (the CPU
    oracle behind the same call shape) - tests/test_adapter_gpu.py.
"""
import types

import torch

from . import guidance_ref, pipeline_ref, unet_ref


class _Fuser:
    def __init__(self):
        self.enabled = True


class OracleUNet:
    """oracle/unet_ref.unet_forward behind the reference's UNet call shape (CPU fp32, torch autograd)"""

    def __init__(self, w, cfg):
        self.w, self.cfg = w, cfg
        self.config = types.SimpleNamespace(in_channels=cfg.in_channels)
        self._fusers = [_Fuser() for _ in range(16)] if cfg.use_gated_attention else []

    def modules(self):
        yield self
        yield from self._fusers

    def __call__(self, sample, t, encoder_hidden_states=None, cross_attention_kwargs=None, **kw):
        ck = cross_attention_kwargs or {}
        save = ck.get("save_attn_to_dict")
        saved = {} if save is not None else None
        gl = ck.get("gligen")
        fuser_on = bool(self._fusers) and all(f.enabled for f in self._fusers)
        keys = [tuple(k) for k in ck["save_keys"]] if ck.get("save_keys") is not None else None
        eps = unet_ref.unet_forward(self.w, self.cfg, sample, t, encoder_hidden_states, gligen=gl, fuser_on=fuser_on,
                                    saved=saved, save_keys=keys)
        if save is not None:
            for k, v in saved.items():
                tok = ck.get("return_token_ca_only")
                if tok is not None:
                    v = v[:, :, :, tok:tok + 1] if isinstance(tok, int) else v[:, :, :, tok]
                if ck.get("return_cond_ca_only"):
                    v = v[v.shape[0] // 2:]
                save[k] = v
        return types.SimpleNamespace(sample=eps)


def enable_fusers(unet, enabled, fuser_types):
    """models/pipelines.py:280-283"""
    for m in unet.modules():
        if isinstance(m, fuser_types):
            m.enabled = enabled


def latent_backward_guidance(sched, unet, cond, index, t, latents, loss, g: pipeline_ref.GuidanceCfg,
                             cross_attention_kwargs=None, trace=None):
    """models/pipelines.py:16-82 through the call shape: saved maps come back in `save_attn_to_dict`, the loss is the
    oracle's compute_ca_lossv3 restatement, the gradient is torch.autograd.grad through whatever `unet` is"""
    it = 0
    if index < g.max_index_step:
        mi = g.max_iter
        if isinstance(mi, list):
            mi = mi[index] if len(mi) > index else mi[-1]
        while float(loss) / g.loss_scale > g.loss_threshold and it < mi:
            saved = {}
            ck = {"save_attn_to_dict": saved, "save_keys": g.keys}
            if cross_attention_kwargs is not None:
                ck.update(cross_attention_kwargs)
            latents = latents.detach().requires_grad_(True)
            with torch.enable_grad():
                unet(latents, t, encoder_hidden_states=cond, return_cross_attention_probs=False,
                     cross_attention_kwargs=ck)
                refs = None
                if g.ref_maps is not None:
                    refs = [[box[index] for box in obj] for obj in g.ref_maps]
                one = {k: v[0].float().cpu() for k, v in saved.items()}
                L = guidance_ref.ca_loss(one, g.bboxes, g.object_positions, g.keys, g.fg_top_p, g.bg_top_p, g.fg_weight,
                                         g.bg_weight, refs, g.word_token_indices, g.ref_ca_loss_weight,
                                         g.ref_word_token_only,
                                         use_ratio_based_loss=g.use_ratio_based_loss) * g.loss_scale
                grad = torch.autograd.grad(L, [latents])[0]
            scale = float((1 - sched.alphas_cumprod[int(t)]) ** 0.5)
            latents = (latents - scale * grad).detach()
            loss = float(L)
            it += 1
            if trace is not None:
                trace.append((index, it, loss))
    return latents, loss, it


@torch.no_grad()
def generate(unet, z0, uncond, cond, steps, guidance_scale=7.5, g=None, gligen=None, gligen_beta=0.3,
             frozen_mask=None, frozen_latents=None, frozen_steps=0, saved_cross_attn_keys=None,
             return_token_ca_only=None, fuser_types=(_Fuser,), trace=None):
    """generate_gligen (models/pipelines.py:324-473) / generate_semantic_guidance (:129-247) over the call shape.
    gligen = dict(boxes [1,30,4], masks [1,30], positive_embeddings [1,30,768]) of the conditional sample."""
    sched = pipeline_ref.DDIM()
    sched.set_timesteps(steps)
    z = z0.clone()
    text = torch.cat([uncond, cond], dim=0)
    main_ck = {"return_cond_ca_only": True, "return_token_ca_only": return_token_ca_only,
               "save_keys": saved_cross_attn_keys, "offload_cross_attn_to_cpu": False}
    guid_ck = {}
    n_ground = int(gligen_beta * len(sched.timesteps))
    if gligen is not None:
        rep = lambda x: torch.cat([x, x], dim=0)
        masks2 = rep(gligen["masks"]).clone()
        masks2[:1] = 0                                                   # pipelines.py:317
        main_ck["gligen"] = dict(boxes=rep(gligen["boxes"]), positive_embeddings=rep(gligen["positive_embeddings"]),
                                 masks=masks2)
        guid_ck["gligen"] = dict(boxes=main_ck["gligen"]["boxes"][:1],
                                 positive_embeddings=main_ck["gligen"]["positive_embeddings"][:1], masks=masks2[:1])
        enable_fusers(unet, True, fuser_types)
    loss = 10000.0
    iters, saved_all = [], []
    for index, t in enumerate(sched.timesteps):
        if gligen is not None and index == n_ground:
            enable_fusers(unet, False, fuser_types)
        if g is not None and g.bboxes:
            z, loss, it = latent_backward_guidance(sched, unet, cond, index, t, z, loss, g, guid_ck or None, trace)
            iters.append(it)
        main_ck["save_attn_to_dict"] = {}
        eps = unet(torch.cat([z, z]), t, encoder_hidden_states=text, cross_attention_kwargs=main_ck).sample
        if saved_cross_attn_keys is not None:
            saved_all.append(main_ck["save_attn_to_dict"])
        del main_ck["save_attn_to_dict"]
        eu, ec = eps.chunk(2)
        z = sched.step(eu + guidance_scale * (ec - eu), t, z)
        if frozen_mask is not None and index < frozen_steps:
            z = frozen_latents[index + 1] * frozen_mask + z * (1.0 - frozen_mask)
    if gligen is not None:
        enable_fusers(unet, False, fuser_types)
    return dict(latents=z, iters=iters, saved=saved_all, loss=loss)
