"""The reference's UNet-operator and AttnProcessor plug-in surface over the B200 engine.

`B200UNetAdapter(net)` is callable exactly like the reference's UNet2DConditionModel inside its pipelines
(models/unet_2d_condition.py:704-719, call sites models/pipelines.py:44,200,427,576):

    out = unet(sample, t, encoder_hidden_states=emb, return_cross_attention_probs=False,
               cross_attention_kwargs={"save_attn_to_dict": d, "save_keys": [...], "return_cond_ca_only": bool,
                                       "return_token_ca_only": int | 1-D index tensor | None,
                                       "offload_cross_attn_to_cpu": bool, "enable_flash_attn": bool,
                                       "gligen": {"boxes", "positive_embeddings", "masks"[, "fuser_attn_kwargs"]}})
    out.sample                                             # eps, same shape / dtype / device as `sample`

and carries `unet.config.in_channels` (utils/latents.py:43,127), `unet.modules()` yielding the GLIGEN fuser handles with
`.enabled` (models/pipelines.py:280-283), `unet.set_attn_processor(...)` / `unet.attn_processors`
(models/unet_2d_condition.py:575-633).  Saved maps follow the AttnProcessor contract
(models/attention_processor.py:463-482): dict[tuple(attn_key)] = [batch, heads, n, tokens], sliced to the token column
and / or the conditional half when asked, on the GPU unless `offload_cross_attn_to_cpu`.

Guidance pass (models/pipelines.py:30-56): when `sample.requires_grad` under enabled grad, the saved maps are returned
as differentiable torch tensors - the forward is the truncated B200 pass with the hand-written tape, and
`torch.autograd.grad(loss, [latents])` runs the hand-written backward chain with d loss / d P injected as dP_extra.  No
torch autograd graph exists inside the network.  `.sample` of such a call is produced lazily by a second, untaped
forward only if it is read (the reference discards it, pipelines.py:44).

Not supported, raised clearly: `attn_process_fn` (arbitrary Python on the attention probabilities cannot run inside the
fused kernel), `attention_mask`, `class_labels`, returning cross-attention probs through the return value
(`return_cross_attention_probs=True` - use save_attn_to_dict).
"""
import types
from typing import Dict, Optional

import torch

from . import ops
from .unet import B200UNet


class FuserHandle:
    """stand-in for a GatedSelfAttentionDense module: the reference toggles `.enabled` on every fuser it finds in
    `unet.modules()` (models/pipelines.py:280-283, models/attention.py:43-53)"""

    def __init__(self, name):
        self.name = name
        self.enabled = True


class UNetOutput:
    """UNet2DConditionOutput-style result: `.sample` (computed lazily for taped guidance calls)"""

    def __init__(self, sample=None, lazy=None):
        self._sample, self._lazy = sample, lazy

    @property
    def sample(self):
        if self._sample is None and self._lazy is not None:
            self._sample = self._lazy()
            self._lazy = None
        return self._sample

    def __getitem__(self, i):
        return (self.sample,)[i]


class _ExtGrad:
    """external d loss / d P holder of one attention key: the backward kernel adds dp_extra to dP"""
    c = None

    def __init__(self, BH, n, dev, ext_ld=80):
        self.dp_extra = torch.zeros(BH, n, ext_ld, device=dev, dtype=torch.float32)


class B200AttnProcessor:
    """AttnProcessor with the reference's signature (models/attention_processor.py:377-393) running ONE attention op of
    a reference-style `attn` module (to_q / to_k / to_v / to_out[0] Linear layers, `.heads`) on the B200 kernels:
    cross-attention through the single-launch xattn_fused kernel (projections + softmax + PV + to_out), self-attention
    through the head-split projection GEMM + tcgen05 attention.  Inference only (no autograd); weights are converted to
    the kernel layout once per module and cached."""

    def __init__(self):
        self._cache = {}

    def _weights(self, attn, dev):
        key = id(attn)
        if key not in self._cache:
            h16 = lambda t: t.detach().to(dev, torch.float16).contiguous()
            wq, wk, wv = attn.to_q.weight, attn.to_k.weight, attn.to_v.weight
            ent = dict(wq=h16(wq), wkv=h16(torch.cat([wk, wv], 0)), wo=h16(attn.to_out[0].weight),
                       bo=attn.to_out[0].bias.detach().to(dev, torch.float32).contiguous()
                       if attn.to_out[0].bias is not None else None)
            if wk.shape[1] == wq.shape[1]:
                ent["wqkv"] = h16(torch.cat([wq, wk, wv], 0))
            self._cache[key] = ent
        return self._cache[key]

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None,
                 return_attntion_probs=False, attn_key=None, attn_process_fn=None, return_cond_ca_only=False,
                 return_token_ca_only=None, offload_cross_attn_to_cpu=False, save_attn_to_dict=None, save_keys=None,
                 enable_flash_attn=True):
        if attn_process_fn is not None:
            raise NotImplementedError("B200AttnProcessor: attn_process_fn (Python on the attention probabilities) cannot "
                                      "run inside the fused cross-attention kernel")
        if attention_mask is not None:
            raise NotImplementedError("B200AttnProcessor: attention_mask is not supported")
        if torch.is_grad_enabled() and hidden_states.requires_grad:
            raise NotImplementedError("B200AttnProcessor is inference-only; the guidance pass runs through "
                                      "B200UNetAdapter (hand-written backward chain)")
        x = hidden_states
        dev, in_dtype = x.device, x.dtype
        Bn, n, C = x.shape
        heads = attn.heads
        d = C // heads
        w = self._weights(attn, dev)
        cross = encoder_hidden_states is not None
        xs = x.reshape(Bn * n, C).to(torch.float16).contiguous()
        scale = float(getattr(attn, "scale", d ** -0.5))
        want = (save_attn_to_dict is not None and (save_keys is None or tuple(attn_key) in save_keys)) \
            or return_attntion_probs
        dp, d16 = ops.round_dp(d), ops.round_d16(d)
        z = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float16)
        probs = None
        if cross:
            ctx = encoder_hidden_states
            T = ctx.shape[1]
            Ta = max(80, (T + 7) // 8 * 8)
            k, v, kt, vt = z(Bn * heads, Ta, dp), z(Bn * heads, Ta, dp), z(Bn * heads, d16, Ta), z(Bn * heads, d16, Ta)
            ops.project_heads2(ctx.reshape(Bn * T, -1).to(torch.float16).contiguous(), w["wkv"], T, heads, d, 1,
                               rm=(None, k, v), tr=(None, kt, vt))
            if T <= 80 and ops.xattn_fused_supported(heads, d, n):
                out, _, _, probs, _ = ops.xattn_fused(xs, w["wq"], k, vt, w["wo"], w["bo"], None, Bn, n, heads, d, T, scale,
                                                      want_probs=want)
            else:
                na = (n + 7) // 8 * 8
                q = z(Bn * heads, na, dp)
                ops.project_heads2(xs, w["wq"], n, heads, d, 0, rm=(q, None, None))
                o, _, probs, _ = ops.xattn_fwd(q, k, vt, Bn, heads, n, T, d, scale, want_probs=want)
                out = ops.linear(o, w["wo"], w["bo"])
        else:
            if want:
                raise NotImplementedError("B200AttnProcessor: self-attention probabilities are never materialised")
            na = (n + 7) // 8 * 8
            q, k, vt = z(Bn * heads, na, dp), z(Bn * heads, na, dp), z(Bn * heads, d16, na)
            ops.project_heads2(xs, w["wqkv"], n, heads, d, 0, rm=(q, k, None), tr=(None, None, vt))
            o = ops.attention_fwd(q, k, vt, Bn, heads, n, n, d, scale)
            out = ops.linear(o, w["wo"], w["bo"])
        hidden = out.view(Bn, n, C).to(in_dtype)
        if getattr(attn, "residual_connection", False):
            hidden = hidden + hidden_states
        rf = getattr(attn, "rescale_output_factor", 1.0)
        if rf != 1.0:
            hidden = hidden / rf
        if want:
            pm = probs.view(Bn, heads, n, -1).to(in_dtype)
            pm = _slice_maps(pm, return_token_ca_only, return_cond_ca_only, offload_cross_attn_to_cpu)
            if save_attn_to_dict is not None and (save_keys is None or tuple(attn_key) in save_keys):
                save_attn_to_dict[tuple(attn_key)] = pm
            if return_attntion_probs:
                return hidden, pm
        return hidden


def _slice_maps(pm, return_token_ca_only, return_cond_ca_only, offload):
    """models/attention_processor.py:466-478"""
    if return_token_ca_only is not None:
        if isinstance(return_token_ca_only, int):
            pm = pm[:, :, :, return_token_ca_only:return_token_ca_only + 1]
        else:
            pm = pm[:, :, :, return_token_ca_only]
    if return_cond_ca_only:
        assert pm.shape[0] % 2 == 0, f"Samples are not in pairs: {pm.shape[0]} samples"
        pm = pm[pm.shape[0] // 2:]
    if offload:
        pm = pm.cpu()
    return pm


class _TapedMaps(torch.autograd.Function):
    """differentiable saved maps of the guidance pass: forward = truncated B200 forward with the tape, backward = the
    hand-written dgrad chain with d loss / d P injected (replaces the autograd graph of models/pipelines.py:44-56)"""

    @staticmethod
    def forward(ctx, sample, adapter, t_dev, kv_fn, keys, objs, fuser_on):
        net = adapter.net
        B, Cz, H, W = sample.shape
        ext = {}
        for k in keys:
            heads, n = adapter._heads_of(k), adapter._tokens_of(k, H, W)
            ext[k] = _ExtGrad(B * heads, n, net.dev)
        saved = {}
        save = dict(keys=list(keys), probs=True, tok=None, out=saved)
        z = sample.detach().to(net.dev, torch.float32).contiguous()
        tape, order = net.guidance_forward(z, t_dev, kv_fn, ext, objs=objs, fuser_on=fuser_on, save=save)
        ctx.adapter, ctx.tape, ctx.ext, ctx.keys, ctx.shape = adapter, tape, ext, list(keys), (B, Cz, H, W)
        ctx.in_dtype, ctx.in_dev = sample.dtype, sample.device
        return tuple(saved[k]["probs"].to(sample.dtype).to(sample.device) for k in keys)

    @staticmethod
    def backward(ctx, *grads):
        net = ctx.adapter.net
        B, Cz, H, W = ctx.shape
        for k, g in zip(ctx.keys, grads):
            e = ctx.ext[k].dp_extra
            e.zero_()
            if g is not None:
                T = g.shape[-1]
                e[:, :, :T] = g.to(net.dev, torch.float32).reshape(e.shape[0], e.shape[1], T) * net.gscale
        g8 = net.guidance_backward(ctx.tape)                        # fp32 [B, HW, 8] NHWC-8, times gscale
        gz = (g8.view(B, H, W, -1)[..., :Cz].permute(0, 3, 1, 2) / net.gscale).contiguous()
        return gz.to(ctx.in_dev, ctx.in_dtype), None, None, None, None, None, None


class B200UNetAdapter:
    def __init__(self, net: B200UNet, fuser_class=None):
        """fuser_class: pass the reference's `models.attention.GatedSelfAttentionDense` when running under the
        reference's own pipelines so that `isinstance(module, GatedSelfAttentionDense)` (models/pipelines.py:282)
        recognises the fuser handles; defaults to FuserHandle."""
        self.net = net
        self.config = types.SimpleNamespace(in_channels=net.cfg.in_channels, out_channels=net.cfg.out_channels,
                                            cross_attention_dim=net.cfg.cross_attention_dim)
        self.dtype, self.device = torch.float32, net.dev
        names = [p + ".fuser" for p, _ in net._attn_layers()] if net.cfg.use_gated_attention else []
        if fuser_class is None:
            self._fusers = [FuserHandle(n) for n in names]
        else:
            # instances of a subclass of the reference's module class, created without running its constructor (no
            # parameters are owned here - the weights live in the engine)
            cls = type("B200" + fuser_class.__name__, (fuser_class,), {})
            self._fusers = []
            for n in names:
                f = cls.__new__(cls)
                f.__dict__.update(name=n, enabled=True)
                self._fusers.append(f)
        self._procs = {p + "." + a + ".processor": B200AttnProcessor() for p, _ in net._attn_layers()
                       for a in ("attn1", "attn2")}
        self._kv_cache = {}
        self._objs_cache = {}
        self._t_dev = {}

    # ---- module-like surface the reference touches
    def modules(self):
        yield self
        for f in self._fusers:
            yield f

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    @property
    def attn_processors(self) -> Dict[str, object]:
        return dict(self._procs)

    def set_attn_processor(self, processor):
        """models/unet_2d_condition.py:596-633.  The engine runs attention inside fused kernels, so only B200AttnProcessor
        instances (or a dict of them keyed like `attn_processors`) are accepted."""
        if isinstance(processor, dict):
            if len(processor) != len(self._procs):
                raise ValueError(f"A dict of processors was passed, but the number of processors {len(processor)} does "
                                 f"not match the number of attention layers: {len(self._procs)}.")
            items = processor
        else:
            items = {k: processor for k in self._procs}
        for k, p in items.items():
            if not isinstance(p, B200AttnProcessor):
                raise NotImplementedError("B200UNetAdapter.set_attn_processor: attention runs inside the fused sm_100a "
                                          f"kernels; only B200AttnProcessor is accepted (got {type(p).__name__})")
            if k not in self._procs:
                raise KeyError(k)
            self._procs[k] = p

    # ---- helpers
    def _heads_of(self, key):
        cfg = self.net.cfg
        if key[0] == "mid":
            return cfg.heads[-1]
        return cfg.heads[key[1]] if key[0] == "down" else list(reversed(cfg.heads))[key[1]]

    def _tokens_of(self, key, H, W):
        nb = len(self.net.cfg.block_out_channels)
        level = {"down": key[1], "mid": nb - 1, "up": nb - 1 - key[1]}[key[0]]
        return (H >> level) * (W >> level)

    def _all_keys(self):
        return [k for _, k in self.net._key_order()]

    def _kv(self, ehs):
        """text K/V of all 16 cross-attention layers, cached per embedding tensor (they do not depend on t or z)"""
        key = (ehs.data_ptr(), tuple(ehs.shape), ehs._version, str(ehs.device))
        if key not in self._kv_cache:
            if len(self._kv_cache) > 8:
                self._kv_cache.clear()
            self._kv_cache[key] = (self.net.set_text(ehs.detach().to(self.net.dev, torch.float32)), ehs)
        return self._kv_cache[key][0]

    def _objs(self, gl):
        key = tuple((gl[k].data_ptr(), gl[k]._version, tuple(gl[k].shape)) for k in ("boxes", "masks", "positive_embeddings"))
        if key not in self._objs_cache:
            if len(self._objs_cache) > 8:
                self._objs_cache.clear()
            self._objs_cache[key] = (self.net.position_net(gl["boxes"].detach(), gl["masks"].detach(),
                                                           gl["positive_embeddings"].detach()), gl)
        return self._objs_cache[key][0]

    def _tvec(self, timestep, B):
        t = float(timestep.item() if torch.is_tensor(timestep) and timestep.numel() == 1 else timestep) \
            if not (torch.is_tensor(timestep) and timestep.numel() > 1) else None
        if B not in self._t_dev:
            self._t_dev[B] = torch.empty(B, device=self.net.dev, dtype=torch.float32)
        if t is None:
            self._t_dev[B].copy_(timestep.to(self.net.dev, torch.float32).reshape(-1).expand(B))
        else:
            self._t_dev[B].fill_(t)
        return self._t_dev[B]

    # ---- the operator
    def __call__(self, sample, timestep, encoder_hidden_states=None, class_labels=None, timestep_cond=None,
                 attention_mask=None, cross_attention_kwargs=None, down_block_additional_residuals=None,
                 mid_block_additional_residual=None, encoder_attention_mask=None, return_dict=True,
                 return_cross_attention_probs=False, **kw):
        if encoder_hidden_states is None:
            raise ValueError("encoder_hidden_states is required")
        for name, v in (("class_labels", class_labels), ("timestep_cond", timestep_cond), ("attention_mask", attention_mask),
                        ("encoder_attention_mask", encoder_attention_mask),
                        ("down_block_additional_residuals", down_block_additional_residuals),
                        ("mid_block_additional_residual", mid_block_additional_residual)):
            if v is not None:
                raise NotImplementedError(f"B200UNetAdapter: `{name}` is not part of the layout-grounded path")
        if return_cross_attention_probs:
            raise NotImplementedError("B200UNetAdapter: pass save_attn_to_dict instead of return_cross_attention_probs")
        ck = dict(cross_attention_kwargs or {})
        if ck.get("attn_process_fn") is not None:
            raise NotImplementedError("B200UNetAdapter: attn_process_fn (Python on the attention probabilities) cannot run "
                                      "inside the fused cross-attention kernel")
        ck.pop("enable_flash_attn", None)            # accepted no-op: the tcgen05 attention kernel is always used
        offload = bool(ck.get("offload_cross_attn_to_cpu", False))
        save_dict = ck.get("save_attn_to_dict")
        save_keys = ck.get("save_keys")
        if save_keys is not None:
            save_keys = [tuple(k) for k in save_keys]
        cond_only = bool(ck.get("return_cond_ca_only", False))
        tok_only = ck.get("return_token_ca_only")
        net = self.net
        B, Cz, H, W = sample.shape
        kv = self._kv(encoder_hidden_states)
        gl = ck.get("gligen")
        objs, fuser_on = None, False
        if gl is not None and net.cfg.use_gated_attention:
            objs = self._objs(gl)
            fuser_on = bool(self._fusers) and all(f.enabled for f in self._fusers)
        t_dev = self._tvec(timestep, B)
        kv_fn = lambda p: kv.slabs[p]

        def full_forward(want_save):
            z = sample.detach().to(net.dev, torch.float32).contiguous()
            kw_save = {}
            if want_save:
                if isinstance(tok_only, int):
                    kw_save = dict(save_keys=save_keys, save_tok=torch.full((B,), tok_only, dtype=torch.int32,
                                                                            device=net.dev))
                else:
                    kw_save = dict(save_keys=save_keys, save_probs=True)
            eps, saved = net.forward(z, t_dev, kv, rep=1, objs=objs, fuser_on=fuser_on, **kw_save)
            return eps.permute(0, 3, 1, 2).to(sample.device, sample.dtype), saved

        taped = torch.is_grad_enabled() and sample.requires_grad and save_dict is not None
        if taped:
            keys = [k for k in self._all_keys() if save_keys is None or k in save_keys]
            maps = _TapedMaps.apply(sample, self, t_dev, kv_fn, keys, objs, fuser_on)
            for k, m in zip(keys, maps):
                save_dict[k] = _slice_maps(m, tok_only, cond_only, offload)
            return UNetOutput(lazy=lambda: full_forward(False)[0])
        eps, saved = full_forward(save_dict is not None)
        if save_dict is not None:
            for k, v in saved.items():
                if v["probs"] is not None:
                    pm = _slice_maps(v["probs"].to(sample.dtype), tok_only, cond_only, offload)
                else:                                    # single token column saved in-kernel: [B, heads, n] -> [..., 1]
                    pm = _slice_maps(v["tok"].to(sample.dtype).unsqueeze(-1), None, cond_only, offload)
                save_dict[k] = pm if pm.device == sample.device or offload else pm.to(sample.device)
        return UNetOutput(sample=eps)
