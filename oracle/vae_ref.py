"""ORACLE (test infrastructure, never on the product path): CPU fp32 restatement of the VAE decode that ends every
generation of the reference (models/pipelines.py:117-127: `vae.decode(latents / 0.18215).sample`, then
(x / 2 + 0.5).clamp(0, 1) -> uint8).

The arithmetic lives in diffusers==0.18.0 (requirements.txt:5; AutoencoderKL, models/vae.py Decoder, UpDecoderBlock2D,
UNetMidBlock2D with one single-head attention, ResnetBlock2D without time embedding) which is absent from /root/reference
and from this image: restated from its published definition, parity against the real wheel is UNPINNED.
SD1.x VAE configuration: latent 4 channels, block_out_channels (128, 256, 512, 512), layers_per_block 2, norm groups 32,
eps 1e-6, scaling factor 0.18215.  Weights use the diffusers state-dict names (post_quant_conv.*, decoder.*).
"""
import math
from dataclasses import dataclass
from typing import Tuple

import torch
import torch.nn.functional as F


@dataclass
class VAEConfig:
    latent_channels: int = 4
    out_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_groups: int = 32
    eps: float = 1e-6
    scaling_factor: float = 0.18215

    @staticmethod
    def tiny():
        return VAEConfig(block_out_channels=(32, 64, 64, 64))


def _gn(x, w, n, cfg):
    return F.group_norm(x, cfg.norm_groups, w[n + ".weight"], w[n + ".bias"], cfg.eps)


def _resnet(x, w, p, cfg):
    h = F.conv2d(F.silu(_gn(x, w, p + ".norm1", cfg)), w[p + ".conv1.weight"], w[p + ".conv1.bias"], padding=1)
    h = F.conv2d(F.silu(_gn(h, w, p + ".norm2", cfg)), w[p + ".conv2.weight"], w[p + ".conv2.bias"], padding=1)
    if p + ".conv_shortcut.weight" in w:
        x = F.conv2d(x, w[p + ".conv_shortcut.weight"], w[p + ".conv_shortcut.bias"])
    return x + h


def _attention(x, w, p, cfg):
    """single-head self-attention over the H*W positions with a GroupNorm in front and a residual connection"""
    B, C, H, W = x.shape
    h = _gn(x, w, p + ".group_norm", cfg).view(B, C, H * W).transpose(1, 2)
    lin = lambda t, n: F.linear(t, w[p + n + ".weight"], w[p + n + ".bias"])
    q, k, v = lin(h, ".to_q"), lin(h, ".to_k"), lin(h, ".to_v")
    pr = torch.softmax(q @ k.transpose(1, 2) * C ** -0.5, dim=-1)
    o = lin(pr @ v, ".to_out.0")
    return x + o.transpose(1, 2).reshape(B, C, H, W)


def decode(w, cfg: VAEConfig, latents):
    """latents [B, 4, h, w] (as they leave the denoising loop) -> float image [B, 3, 8h, 8w] (AutoencoderKL.decode)"""
    z = latents / cfg.scaling_factor
    z = F.conv2d(z, w["post_quant_conv.weight"], w["post_quant_conv.bias"])
    h = F.conv2d(z, w["decoder.conv_in.weight"], w["decoder.conv_in.bias"], padding=1)
    h = _resnet(h, w, "decoder.mid_block.resnets.0", cfg)
    h = _attention(h, w, "decoder.mid_block.attentions.0", cfg)
    h = _resnet(h, w, "decoder.mid_block.resnets.1", cfg)
    nb = len(cfg.block_out_channels)
    for i in range(nb):
        for j in range(cfg.layers_per_block + 1):
            h = _resnet(h, w, f"decoder.up_blocks.{i}.resnets.{j}", cfg)
        if i < nb - 1:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            p = f"decoder.up_blocks.{i}.upsamplers.0.conv"
            h = F.conv2d(h, w[p + ".weight"], w[p + ".bias"], padding=1)
    h = F.silu(_gn(h, w, "decoder.conv_norm_out", cfg))
    return F.conv2d(h, w["decoder.conv_out.weight"], w["decoder.conv_out.bias"], padding=1)


def to_uint8(img):
    """models/pipelines.py:124-126"""
    x = (img / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1).numpy()
    return (x * 255).round().astype("uint8")


def shapes(cfg: VAEConfig):
    out = []
    conv = lambda n, co, ci, k: out.extend([(n + ".weight", (co, ci, k, k)), (n + ".bias", (co,))])
    norm = lambda n, c: out.extend([(n + ".weight", (c,)), (n + ".bias", (c,))])
    lin = lambda n, o, i: out.extend([(n + ".weight", (o, i)), (n + ".bias", (o,))])

    def res(n, ci, co):
        norm(n + ".norm1", ci); conv(n + ".conv1", co, ci, 3); norm(n + ".norm2", co); conv(n + ".conv2", co, co, 3)
        if ci != co:
            conv(n + ".conv_shortcut", co, ci, 1)
    rc = list(reversed(cfg.block_out_channels))
    conv("post_quant_conv", cfg.latent_channels, cfg.latent_channels, 1)
    conv("decoder.conv_in", rc[0], cfg.latent_channels, 3)
    res("decoder.mid_block.resnets.0", rc[0], rc[0])
    a = "decoder.mid_block.attentions.0"
    norm(a + ".group_norm", rc[0])
    for n in (".to_q", ".to_k", ".to_v", ".to_out.0"):
        lin(a + n, rc[0], rc[0])
    res("decoder.mid_block.resnets.1", rc[0], rc[0])
    ch = rc[0]
    for i, co in enumerate(rc):
        for j in range(cfg.layers_per_block + 1):
            res(f"decoder.up_blocks.{i}.resnets.{j}", ch, co)
            ch = co
        if i < len(rc) - 1:
            conv(f"decoder.up_blocks.{i}.upsamplers.0.conv", ch, ch, 3)
    norm("decoder.conv_norm_out", ch)
    conv("decoder.conv_out", cfg.out_channels, ch, 3)
    return out


def make_weights(cfg: VAEConfig, seed=0):
    """seeded synthetic decoder weights of the real shapes (no checkpoints offline)"""
    g = torch.Generator().manual_seed(seed)
    w = {}
    for name, shape in shapes(cfg):
        if "norm" in name.split(".")[-2] and name.endswith(".weight"):
            w[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith(".bias"):
            w[name] = 0.05 * torch.randn(shape, generator=g)
        else:
            scale = 1.0 / math.sqrt(math.prod(shape[1:]))
            if name.endswith(("conv2.weight", "to_out.0.weight")):
                scale *= 0.5
            if name.endswith(("to_q.weight", "to_k.weight")):
                scale *= 2.0
            w[name] = scale * torch.randn(shape, generator=g)
    return w
