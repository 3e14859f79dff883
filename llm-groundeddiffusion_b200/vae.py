"""VAE decode on the B200 kernels: `decode(latents)` of models/pipelines.py:117-127
- AutoencoderKL.decode(latents / 0.18215).sample -> (x / 2 + 0.5).clamp(0, 1) -> uint8 [B, 8h, 8w, 3] - called once per
per-box generation and once per overall generation (models/pipelines.py:233,461,589-591).

The decoder (diffusers 0.18 models/vae.py: conv_in, mid block = resnet / single-head attention / resnet, four
UpDecoderBlock2D of three resnets + nearest-x2 upsample conv, GroupNorm + SiLU + conv_out) is sequenced over the same
hand-written sm_100a entry points as the UNet: implicit-GEMM convolutions, GroupNorm(+SiLU), the linear GEMM; the one
single-head 512-wide attention (4096 tokens at 512x512) is two GEMMs around a row-softmax kernel because its head
dimension does not fit the fused attention tiles.  fp16 activations NHWC, fp32 accumulation; weights in diffusers
state-dict names (`post_quant_conv.*`, `decoder.*`).  to_v's bias is folded into to_out's (softmax rows sum to 1).
"""
import ctypes
from dataclasses import dataclass
from typing import Dict, Tuple

import torch

from . import ops
from ._lib import check, cur_stream, lib, ptr
from .unet import B200UNet, TAPS_1x1, UNetConfig, gemm

_i, _f = ctypes.c_int, ctypes.c_float


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


class B200VAEDecoder(B200UNet):
    def __init__(self, cfg: VAEConfig, weights: Dict[str, torch.Tensor], device="cuda:0"):
        self.vcfg = cfg
        sd = {}
        for k, v in weights.items():
            if k.startswith("decoder."):
                sd[k[len("decoder."):]] = v
        # conv_out: pad the 3 output channels to 8 (zero rows) so the epilogue stores whole vectors
        co = sd["conv_out.weight"]
        pad_w = co.new_zeros(8, *co.shape[1:])
        pad_w[:co.shape[0]] = co
        pad_b = co.new_zeros(8)
        pad_b[:co.shape[0]] = sd["conv_out.bias"]
        sd["conv_out.weight"], sd["conv_out.bias"] = pad_w, pad_b
        # fold to_v's bias into to_out's: P (V + 1 b_v^T) Wo^T + b_o = P V Wo^T + (Wo b_v + b_o)
        a = "mid_block.attentions.0"
        sd = dict(sd)
        sd[a + ".to_out.0.bias"] = sd[a + ".to_out.0.bias"] + sd[a + ".to_out.0.weight"] @ sd[a + ".to_v.bias"]
        super().__init__(UNetConfig(in_channels=cfg.latent_channels, norm_groups=cfg.norm_groups, norm_eps=cfg.eps), sd,
                         device)
        f32 = lambda t: t.detach().to(self.dev, torch.float32).contiguous()
        self.pq_w = f32(weights["post_quant_conv.weight"].reshape(cfg.latent_channels, cfg.latent_channels))
        self.pq_b = f32(weights["post_quant_conv.bias"])
        if cfg.latent_channels != 4:
            raise NotImplementedError("B200VAEDecoder: 4 latent channels")

    def _res(self, x, p):
        eps = self.vcfg.eps
        h = self.conv(self.group_norm(x, p + ".norm1", eps, True), p + ".conv1")
        h = self.group_norm(h, p + ".norm2", eps, True)
        sc = x
        if (p + ".conv_shortcut.w") in self.w:
            sc = self.conv(x, p + ".conv_shortcut", taps=TAPS_1x1)
        return self.conv(h, p + ".conv2", residual=sc)

    def _attention(self, x, p):
        B, H, W, C = x.shape
        n = H * W
        h = self.group_norm(x, p + ".group_norm", self.vcfg.eps, False).view(B * n, C)
        q = self.linear(h, p + ".to_q")
        k = self.linear(h, p + ".to_k")
        # V^T [C, n] per image directly from the GEMM: (Wv h^T) = linear(A = Wv, W = h); its bias is folded into to_out
        wv = self.w[p + ".to_v.w"]
        o = torch.empty(B * n, C, device=self.dev, dtype=torch.float16)
        scores = torch.empty(n, n, device=self.dev, dtype=torch.float32)
        probs = torch.empty(n, n, device=self.dev, dtype=torch.float16)
        vt = torch.empty(C, n, device=self.dev, dtype=torch.float16)
        for b in range(B):
            hb, qb, kb = h[b * n:(b + 1) * n], q[b * n:(b + 1) * n], k[b * n:(b + 1) * n]
            gemm(wv, (1, 1, C, C, C), hb, n, 1, (1, 1, C), TAPS_1x1, out=vt, ldo=n)
            gemm(qb, (1, 1, n, C, C), kb, n, 1, (1, 1, n), TAPS_1x1, out_f32=scores, ldo32=n, alpha=C ** -0.5)
            check(lib().b200lmd_softmax_rows(ptr(scores), ptr(probs), ctypes.c_longlong(n), _i(n), cur_stream()))
            gemm(probs, (1, 1, n, n, n), vt, C, 1, (1, 1, n), TAPS_1x1, out=o[b * n:(b + 1) * n], ldo=C)
        y = self.linear(o, p + ".to_out.0", residual=x.view(B * n, C))
        return y.view(B, H, W, C)

    def decode_raw(self, latents):
        """latents fp32 [B, 4, h, w] (any device) -> decoder output fp32 NHWC [B, 8h, 8w, 8] (channels 0..2 valid)"""
        cfg = self.vcfg
        self.tape = None
        z = latents.to(self.dev, torch.float32).contiguous()
        B, Cz, H, W = z.shape
        x = torch.empty(B, H, W, 8, device=self.dev, dtype=torch.float16)
        check(lib().b200lmd_vae_prepare_latents(ptr(z), ptr(self.pq_w), ptr(self.pq_b), ptr(x), _i(B), _i(H * W),
                                                _f(1.0 / cfg.scaling_factor), cur_stream()))
        h = self.conv(x, "conv_in")
        h = self._res(h, "mid_block.resnets.0")
        h = self._attention(h, "mid_block.attentions.0")
        h = self._res(h, "mid_block.resnets.1")
        nb = len(cfg.block_out_channels)
        for i in range(nb):
            for j in range(cfg.layers_per_block + 1):
                h = self._res(h, f"up_blocks.{i}.resnets.{j}")
            if i < nb - 1:
                h = self.upsample_conv(h, f"up_blocks.{i}.upsamplers.0.conv")
        h = self.group_norm(h, "conv_norm_out", cfg.eps, True)
        return self.conv(h, "conv_out", out_f32=True)

    def decode(self, latents, chunk=4):
        """-> uint8 [B, 8h, 8w, 3] on the device (models/pipelines.py:117-127); images are decoded `chunk` at a time to
        bound the 512x512x128-channel activations"""
        outs = []
        for b0 in range(0, latents.shape[0], chunk):
            raw = self.decode_raw(latents[b0:b0 + chunk])
            B, H, W, ld = raw.shape
            img = torch.empty(B, H, W, 3, device=self.dev, dtype=torch.uint8)
            check(lib().b200lmd_vae_to_uint8(ptr(raw), _i(ld), ptr(img), ctypes.c_longlong(B * H * W), cur_stream()))
            outs.append(img)
        return torch.cat(outs, 0)
