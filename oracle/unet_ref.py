"""ORACLE (test infrastructure, never on the product path): CPU fp32 restatement of the reference UNet forward.

A single functional pass over a flat {name: tensor} weight dict that uses the diffusers state-dict names, so the same
weights feed the reference modules (validation, build container only), this oracle (travels to the GPU box) and the
CUDA engine.  Restates, citing the reference:
  UNet2DConditionModel.forward          models/unet_2d_condition.py:704-980
  CrossAttnDown/Mid/CrossAttnUp/Up/Down models/unet_2d_blocks.py:370-451, 245-278, 627-709, 762-793, 506-537
  Transformer2DModel.forward            models/transformer_2d.py:216-367
  BasicTransformerBlock / FeedForward   models/attention.py:156-237, 286-335
  GatedSelfAttentionDense / PositionNet models/attention.py:43-53, models/unet_2d_condition.py:63-114
  Attention + AttnProcessor             models/attention_processor.py:201-233, 377-483
  ResnetBlock2D/Downsample2D/Upsample2D/Timesteps/TimestepEmbedding: diffusers==0.18.0 (requirements.txt:5; absent
  from /root/reference and from this image) - restated from its published definitions; parity for those pieces is
  pinned only through oracle/shim, i.e. "parity unpinned" against the real wheel.
"""
import math
from dataclasses import dataclass, field
from typing import Optional, Tuple

import torch
import torch.nn.functional as F


@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    heads: Tuple[int, ...] = (8, 8, 8, 8)          # diffusers' `attention_head_dim` is the head COUNT for SD1.x
    cross_attention_dim: int = 768
    norm_groups: int = 32
    norm_eps: float = 1e-5
    use_linear_projection: bool = False
    use_gated_attention: bool = False
    down_attn: Tuple[bool, ...] = (True, True, True, False)
    up_attn: Tuple[bool, ...] = (False, True, True, True)

    @staticmethod
    def sd15(gligen=False):
        return UNetConfig(use_gated_attention=gligen)

    @staticmethod
    def sd21():
        return UNetConfig(heads=(5, 10, 20, 20), cross_attention_dim=1024, use_linear_projection=True)

    @staticmethod
    def tiny(gligen=False):
        """same topology, 128..512 channels: CPU-fast parity config"""
        return UNetConfig(block_out_channels=(128, 256, 512, 512), cross_attention_dim=768, use_gated_attention=gligen)

    def to_reference_kwargs(self):
        return dict(in_channels=self.in_channels, out_channels=self.out_channels,
                    block_out_channels=self.block_out_channels, layers_per_block=self.layers_per_block,
                    attention_head_dim=self.heads, cross_attention_dim=self.cross_attention_dim,
                    norm_num_groups=self.norm_groups, norm_eps=self.norm_eps,
                    use_linear_projection=self.use_linear_projection, use_gated_attention=self.use_gated_attention)


def timestep_embedding(t, dim):
    """diffusers embeddings.get_timestep_embedding with flip_sin_to_cos=True, freq_shift=0 -> [cos, sin]"""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    ang = t[:, None].float() * freqs[None]
    return torch.cat([ang.cos(), ang.sin()], dim=-1)


def _gn(x, w, name, groups, eps):
    return F.group_norm(x, groups, w[name + ".weight"], w[name + ".bias"], eps)


def _lin(x, w, name, bias=True):
    return F.linear(x, w[name + ".weight"], w.get(name + ".bias") if bias else None)


def resnet(x, temb, w, p, cfg):
    h = F.conv2d(F.silu(_gn(x, w, p + ".norm1", cfg.norm_groups, cfg.norm_eps)), w[p + ".conv1.weight"],
                 w[p + ".conv1.bias"], padding=1)
    h = h + _lin(F.silu(temb), w, p + ".time_emb_proj")[:, :, None, None]
    h = F.conv2d(F.silu(_gn(h, w, p + ".norm2", cfg.norm_groups, cfg.norm_eps)), w[p + ".conv2.weight"],
                 w[p + ".conv2.bias"], padding=1)
    if p + ".conv_shortcut.weight" in w:
        x = F.conv2d(x, w[p + ".conv_shortcut.weight"], w[p + ".conv_shortcut.bias"])
    return x + h


def attention(x, ctx, w, p, heads, want_probs=False):
    """softmax(scale q k^T) v, then to_out (models/attention_processor.py:426-453); probs [B, heads, n, T]"""
    B, n, C = x.shape
    src = x if ctx is None else ctx
    q = _lin(x, w, p + ".to_q", bias=False)
    k = _lin(src, w, p + ".to_k", bias=False)
    v = _lin(src, w, p + ".to_v", bias=False)
    d = C // heads
    q = q.view(B, n, heads, d).transpose(1, 2)
    k = k.view(B, -1, heads, d).transpose(1, 2)
    v = v.view(B, -1, heads, d).transpose(1, 2)
    probs = torch.softmax((q @ k.transpose(-1, -2)) * d ** -0.5, dim=-1)
    o = (probs @ v).transpose(1, 2).reshape(B, n, C)
    o = _lin(o, w, p + ".to_out.0")
    return (o, probs) if want_probs else o


def feed_forward(x, w, p):
    h, gate = _lin(x, w, p + ".net.0.proj").chunk(2, dim=-1)
    return _lin(h * F.gelu(gate), w, p + ".net.2")


def _ln(x, w, name):
    return F.layer_norm(x, (x.shape[-1],), w[name + ".weight"], w[name + ".bias"], 1e-5)


def fuser(x, objs, w, p, heads):
    """GatedSelfAttentionDense (models/attention.py:43-53)"""
    n = x.shape[1]
    o = _lin(objs, w, p + ".linear")
    a = attention(_ln(torch.cat([x, o], dim=1), w, p + ".norm1"), None, w, p + ".attn", heads)[:, :n]
    x = x + torch.tanh(w[p + ".alpha_attn"]) * a
    x = x + torch.tanh(w[p + ".alpha_dense"]) * feed_forward(_ln(x, w, p + ".norm2"), w, p + ".ff")
    return x


def transformer(x, ctx, w, p, heads, cfg, key, saved, save_keys, objs, fuser_on):
    B, C, H, W = x.shape
    res = x
    h = _gn(x, w, p + ".norm", cfg.norm_groups, 1e-6)
    if not cfg.use_linear_projection:
        h = F.conv2d(h, w[p + ".proj_in.weight"], w[p + ".proj_in.bias"])
        h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
    else:
        h = _lin(h.permute(0, 2, 3, 1).reshape(B, H * W, C), w, p + ".proj_in")
    b = p + ".transformer_blocks.0"
    h = h + attention(_ln(h, w, b + ".norm1"), None, w, b + ".attn1", heads)
    if cfg.use_gated_attention and objs is not None and fuser_on:
        h = fuser(h, objs, w, b + ".fuser", heads)
    want = saved is not None and (save_keys is None or key in save_keys)
    a = attention(_ln(h, w, b + ".norm2"), ctx, w, b + ".attn2", heads, want_probs=want)
    if want:
        a, probs = a
        saved[key] = probs
    h = h + a
    h = h + feed_forward(_ln(h, w, b + ".norm3"), w, b + ".ff")
    if not cfg.use_linear_projection:
        h = h.reshape(B, H, W, C).permute(0, 3, 1, 2)
        h = F.conv2d(h, w[p + ".proj_out.weight"], w[p + ".proj_out.bias"])
    else:
        h = _lin(h, w, p + ".proj_out").reshape(B, H, W, C).permute(0, 3, 1, 2)
    return h + res


def position_net(boxes, masks, emb, w):
    """PositionNet + FourierEmbedder(num_freqs=8, temperature=100) (models/unet_2d_condition.py:63-114)"""
    m = masks.unsqueeze(-1)
    freqs = 100.0 ** (torch.arange(8, dtype=torch.float32) / 8)
    ang = boxes.unsqueeze(-1) * freqs                       # [B, N, 4, 8]
    four = torch.stack([ang.sin(), ang.cos()], dim=-1)       # [B, N, 4, 8, 2]
    four = four.permute(0, 1, 3, 4, 2).reshape(*boxes.shape[:2], -1)
    emb = emb * m + (1 - m) * w["position_net.null_positive_feature"].view(1, 1, -1)
    four = four * m + (1 - m) * w["position_net.null_position_feature"].view(1, 1, -1)
    h = torch.cat([emb, four], dim=-1)
    h = F.silu(_lin(h, w, "position_net.linears.0"))
    h = F.silu(_lin(h, w, "position_net.linears.2"))
    return _lin(h, w, "position_net.linears.4")


def unet_forward(w, cfg: UNetConfig, sample, t, ctx, gligen=None, fuser_on=True, saved=None, save_keys=None,
                 stop_after_key=None):
    """sample [B,4,H,W] fp32, t scalar or [B], ctx [B,T,ctx_dim].  gligen = dict(boxes [B,30,4], masks [B,30],
    positive_embeddings [B,30,768]) or None.  `saved`: dict filled with probs [B,heads,n,T] for keys in save_keys
    (None = all), keys are the reference's tuples, e.g. ("up", 1, 2, 0) (models/pipelines.py:12-14).
    Returns eps [B,4,H,W]."""
    B = sample.shape[0]
    t = torch.as_tensor(t, dtype=torch.float32).reshape(-1).expand(B)
    c0 = cfg.block_out_channels[0]
    temb = _lin(F.silu(_lin(timestep_embedding(t, c0), w, "time_embedding.linear_1")), w, "time_embedding.linear_2")
    h = F.conv2d(sample, w["conv_in.weight"], w["conv_in.bias"], padding=1)
    objs = None
    if gligen is not None:
        objs = position_net(gligen["boxes"], gligen["masks"], gligen["positive_embeddings"], w)
    skips = [h]
    nb = len(cfg.block_out_channels)
    for i in range(nb):
        for j in range(cfg.layers_per_block):
            h = resnet(h, temb, w, f"down_blocks.{i}.resnets.{j}", cfg)
            if cfg.down_attn[i]:
                h = transformer(h, ctx, w, f"down_blocks.{i}.attentions.{j}", cfg.heads[i], cfg, ("down", i, j, 0),
                                saved, save_keys, objs, fuser_on)
            skips.append(h)
        if i < nb - 1:
            p = f"down_blocks.{i}.downsamplers.0.conv"
            h = F.conv2d(h, w[p + ".weight"], w[p + ".bias"], stride=2, padding=1)
            skips.append(h)
    h = resnet(h, temb, w, "mid_block.resnets.0", cfg)
    h = transformer(h, ctx, w, "mid_block.attentions.0", cfg.heads[-1], cfg, ("mid", 0, 0, 0), saved, save_keys, objs,
                    fuser_on)
    h = resnet(h, temb, w, "mid_block.resnets.1", cfg)
    rheads = list(reversed(cfg.heads))
    for i in range(nb):
        for j in range(cfg.layers_per_block + 1):
            h = torch.cat([h, skips.pop()], dim=1)
            h = resnet(h, temb, w, f"up_blocks.{i}.resnets.{j}", cfg)
            if cfg.up_attn[i]:
                h = transformer(h, ctx, w, f"up_blocks.{i}.attentions.{j}", rheads[i], cfg, ("up", i, j, 0), saved,
                                save_keys, objs, fuser_on)
        if i < nb - 1:
            p = f"up_blocks.{i}.upsamplers.0.conv"
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, w[p + ".weight"], w[p + ".bias"], padding=1)
    h = F.silu(_gn(h, w, "conv_norm_out", cfg.norm_groups, cfg.norm_eps))
    return F.conv2d(h, w["conv_out.weight"], w["conv_out.bias"], padding=1)


# --------------------------------------------------------------------------------------------- synthetic weights
def _shapes(cfg: UNetConfig):
    """(name, shape) for every parameter of the architecture, in diffusers state-dict naming"""
    out = []
    C = cfg.block_out_channels
    T = C[0] * 4
    X = cfg.cross_attention_dim

    def conv(n, co, ci, k):
        out.extend([(n + ".weight", (co, ci, k, k)), (n + ".bias", (co,))])

    def lin(n, o, i, bias=True):
        out.append((n + ".weight", (o, i)))
        if bias:
            out.append((n + ".bias", (o,)))

    def norm(n, c):
        out.extend([(n + ".weight", (c,)), (n + ".bias", (c,))])

    def res(n, ci, co):
        norm(n + ".norm1", ci); conv(n + ".conv1", co, ci, 3); lin(n + ".time_emb_proj", co, T)
        norm(n + ".norm2", co); conv(n + ".conv2", co, co, 3)
        if ci != co:
            conv(n + ".conv_shortcut", co, ci, 1)

    def attn(n, c, kv):
        lin(n + ".to_q", c, c, False); lin(n + ".to_k", c, kv, False); lin(n + ".to_v", c, kv, False)
        lin(n + ".to_out.0", c, c)

    def ff(n, c):
        lin(n + ".net.0.proj", 8 * c, c); lin(n + ".net.2", c, 4 * c)

    def tr(n, c):
        norm(n + ".norm", c)
        if cfg.use_linear_projection:
            lin(n + ".proj_in", c, c); lin(n + ".proj_out", c, c)
        else:
            conv(n + ".proj_in", c, c, 1); conv(n + ".proj_out", c, c, 1)
        b = n + ".transformer_blocks.0"
        norm(b + ".norm1", c); attn(b + ".attn1", c, c)
        norm(b + ".norm2", c); attn(b + ".attn2", c, X)
        norm(b + ".norm3", c); ff(b + ".ff", c)
        if cfg.use_gated_attention:
            f = b + ".fuser"
            lin(f + ".linear", c, X); attn(f + ".attn", c, c); ff(f + ".ff", c)
            norm(f + ".norm1", c); norm(f + ".norm2", c)
            out.extend([(f + ".alpha_attn", ()), (f + ".alpha_dense", ())])

    conv("conv_in", C[0], cfg.in_channels, 3)
    lin("time_embedding.linear_1", T, C[0]); lin("time_embedding.linear_2", T, T)
    nb = len(C)
    ch = C[0]
    skip_ch = [C[0]]
    for i in range(nb):
        for j in range(cfg.layers_per_block):
            res(f"down_blocks.{i}.resnets.{j}", ch, C[i]); ch = C[i]
            if cfg.down_attn[i]:
                tr(f"down_blocks.{i}.attentions.{j}", ch)
            skip_ch.append(ch)
        if i < nb - 1:
            conv(f"down_blocks.{i}.downsamplers.0.conv", ch, ch, 3)
            skip_ch.append(ch)
    res("mid_block.resnets.0", ch, ch); tr("mid_block.attentions.0", ch); res("mid_block.resnets.1", ch, ch)
    rc = list(reversed(C))
    for i in range(nb):
        for j in range(cfg.layers_per_block + 1):
            res(f"up_blocks.{i}.resnets.{j}", ch + skip_ch.pop(), rc[i]); ch = rc[i]
            if cfg.up_attn[i]:
                tr(f"up_blocks.{i}.attentions.{j}", ch)
        if i < nb - 1:
            conv(f"up_blocks.{i}.upsamplers.0.conv", ch, ch, 3)
    norm("conv_norm_out", ch); conv("conv_out", cfg.out_channels, ch, 3)
    if cfg.use_gated_attention:
        lin("position_net.linears.0", 512, 768 + 64); lin("position_net.linears.2", 512, 512)
        lin("position_net.linears.4", X, 512)
        out.extend([("position_net.null_positive_feature", (768,)), ("position_net.null_position_feature", (64,))])
    return out


def make_weights(cfg: UNetConfig, seed=0, qk_gain=3.0):
    """Seeded synthetic weights with the architecture's real shapes (no SD checkpoints exist offline, SURVEY.md
    fact 3).  Scales are fan-in normalised so activations stay O(1) through ~60 layers; cross-attention q/k get an
    extra gain so softmax over the 77 text tokens is not uniform (otherwise the guidance loss has no signal)."""
    g = torch.Generator().manual_seed(seed)
    w = {}
    for name, shape in _shapes(cfg):
        if name.endswith("alpha_attn") or name.endswith("alpha_dense"):
            w[name] = torch.tensor(0.6 if name.endswith("alpha_attn") else -0.4)
        elif "norm" in name.split(".")[-2] and name.endswith(".weight") and len(shape) == 1:
            w[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias") or name.startswith("position_net.null"):
            w[name] = 0.05 * torch.randn(shape, generator=g)
        else:
            fan_in = math.prod(shape[1:])
            scale = 1.0 / math.sqrt(fan_in)
            if ".attn2.to_q" in name or ".attn2.to_k" in name:
                scale *= qk_gain
            if name.endswith("to_out.0.weight") or name.endswith("net.2.weight") or name.endswith("conv2.weight") \
                    or name.endswith("proj_out.weight"):
                scale *= 0.5      # residual branches: keep the trunk from blowing up
            w[name] = scale * torch.randn(shape, generator=g)
    return w
