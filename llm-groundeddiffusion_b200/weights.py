"""Parameter inventory of the UNet (diffusers state-dict names) and synthetic weights of the real shapes.

No SD / GLIGEN checkpoints exist offline; bench.py and smoke() therefore run on seeded random
weights with the true SD1.5 / SD1.4+GLIGEN / SD2.1 layer shapes.  Real checkpoints load through the same names
(`B200UNet(cfg, state_dict)`)."""
import math

import torch


def parameter_shapes(cfg):
    out = []
    C = cfg.block_out_channels
    T = C[0] * 4
    X = cfg.cross_attention_dim
    conv = lambda n, co, ci, k: out.extend([(n + ".weight", (co, ci, k, k)), (n + ".bias", (co,))])
    norm = lambda n, c: out.extend([(n + ".weight", (c,)), (n + ".bias", (c,))])

    def lin(n, o, i, bias=True):
        out.append((n + ".weight", (o, i)))
        if bias:
            out.append((n + ".bias", (o,)))

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
    ch, skip = C[0], [C[0]]
    for i in range(nb):
        for j in range(cfg.layers_per_block):
            res(f"down_blocks.{i}.resnets.{j}", ch, C[i]); ch = C[i]
            if cfg.down_attn[i]:
                tr(f"down_blocks.{i}.attentions.{j}", ch)
            skip.append(ch)
        if i < nb - 1:
            conv(f"down_blocks.{i}.downsamplers.0.conv", ch, ch, 3); skip.append(ch)
    res("mid_block.resnets.0", ch, ch); tr("mid_block.attentions.0", ch); res("mid_block.resnets.1", ch, ch)
    rc = list(reversed(C))
    for i in range(nb):
        for j in range(cfg.layers_per_block + 1):
            res(f"up_blocks.{i}.resnets.{j}", ch + skip.pop(), rc[i]); ch = rc[i]
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


def synthetic_weights(cfg, seed=0, device="cpu", qk_gain=3.0):
    """fan-in normalised random weights (residual branches damped, cross-attention q/k sharpened so the 77-token
    softmax is not uniform); generated on `device`."""
    g = torch.Generator(device=device).manual_seed(seed)
    w = {}
    rn = lambda shape: torch.randn(shape, generator=g, device=device)
    for name, shape in parameter_shapes(cfg):
        if name.endswith("alpha_attn") or name.endswith("alpha_dense"):
            w[name] = torch.tensor(0.6 if name.endswith("alpha_attn") else -0.4, device=device)
        elif "norm" in name.split(".")[-2] and name.endswith(".weight") and len(shape) == 1:
            w[name] = 1.0 + 0.1 * rn(shape)
        elif name.endswith(".bias") or name.startswith("position_net.null"):
            w[name] = 0.05 * rn(shape)
        else:
            scale = 1.0 / math.sqrt(math.prod(shape[1:]))
            if ".attn2.to_q" in name or ".attn2.to_k" in name:
                scale *= qk_gain
            if name.endswith(("to_out.0.weight", "net.2.weight", "conv2.weight", "proj_out.weight")):
                scale *= 0.5
            w[name] = scale * rn(shape)
    return w


def vae_parameter_shapes(cfg):
    """diffusers-0.18 AutoencoderKL decoder side (post_quant_conv.*, decoder.*) for a lgd_b200.vae.VAEConfig"""
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


def synthetic_vae_weights(cfg, seed=0, device="cpu"):
    """seeded decoder weights of the real shapes (no VAE checkpoint exists offline)"""
    g = torch.Generator(device=device).manual_seed(seed)
    rn = lambda shape: torch.randn(shape, generator=g, device=device)
    w = {}
    for name, shape in vae_parameter_shapes(cfg):
        if "norm" in name.split(".")[-2] and name.endswith(".weight"):
            w[name] = 1.0 + 0.1 * rn(shape)
        elif name.endswith(".bias"):
            w[name] = 0.05 * rn(shape)
        else:
            scale = 1.0 / math.sqrt(math.prod(shape[1:]))
            if name.endswith(("conv2.weight", "to_out.0.weight")):
                scale *= 0.5
            w[name] = scale * rn(shape)
    return w
