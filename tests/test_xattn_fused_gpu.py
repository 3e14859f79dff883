"""GPU parity of the single-launch cross-attention op (xattn_fused_kernel: x Wq^T -> QK^T -> softmax -> PV -> Wo^T +
bias + residual, with the guidance loss) against the three-launch path it replaces and against fp32 PyTorch."""
import pytest
import torch

pytestmark = pytest.mark.gpu
KEY = ("up", 1, 0, 0)


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


@pytest.mark.parametrize("B,d,n,with_loss", [(8, 160, 256, True), (2, 160, 256, False), (3, 64, 256, True),
                                             (2, 80, 1024, False), (1, 160, 256, True)])
def test_fused_matches_unfused_and_torch(cuda, B, d, n, with_loss):
    from lgd_b200 import guidance as G, ops
    from test_xattn_loss_gpu import _layouts
    heads, nk = 8, 77
    C = heads * d
    g = torch.Generator(device="cpu").manual_seed(n + d + B)
    x = torch.randn(B * n, C, generator=g).half().to(cuda)
    res = torch.randn(B * n, C, generator=g).half().to(cuda)
    ctx = torch.randn(B * nk, 768, generator=g).half().to(cuda)
    wq = (torch.randn(C, C, generator=g) * 3 / C ** 0.5).half().to(cuda)
    wkv = (torch.randn(2 * C, 768, generator=g) * 2 / 768 ** 0.5).half().to(cuda)
    wo = (torch.randn(C, C, generator=g) / C ** 0.5).half().to(cuda)
    bo = torch.randn(C, generator=g).to(cuda)
    dp, d16 = ops.round_dp(d), ops.round_d16(d)
    z = lambda *s: torch.zeros(*s, device=cuda, dtype=torch.float16)
    k, v, kt, vt = z(B * heads, 80, dp), z(B * heads, 80, dp), z(B * heads, d16, 80), z(B * heads, d16, 80)
    ops.project_heads2(ctx, wkv, nk, heads, d, 1, rm=(None, k, v), tr=(None, kt, vt))
    scale = d ** -0.5
    kl = kl2 = None
    if with_loss:
        samples = _layouts(B, n + d, heads, n, True)
        params = G.LossParams(loss_scale=5.0, fg_weight=1.0, bg_weight=4.0, ref_ca_loss_weight=2.0,
                              ref_word_token_only=True, use_ref=True)
        slot_tok, slot_of = G.assign_slots(samples, params)
        slot_dev = torch.from_numpy(slot_tok).to(cuda)
        kl = G.KeyLoss(samples, slot_dev, slot_of, KEY, n, heads, 4, params, cuda, gscale=64.0)
        kl2 = G.KeyLoss(samples, slot_dev, slot_of, KEY, n, heads, 4, params, cuda, gscale=64.0)
    save_tok = torch.tensor([3] * B, dtype=torch.int32, device=cuda)
    out, q_f, lse_f, probs_f, ptok_f = ops.xattn_fused(x, wq, k, vt, wo, bo, res, B, n, heads, d, nk, scale, loss=kl,
                                                       want_probs=True, save_tok=save_tok, want_q=True)
    torch.cuda.synchronize()
    # three-launch path
    q = z(B * heads, n, dp)
    ops.project_heads2(x, wq, n, heads, d, 0, rm=(q, None, None))
    o, lse, probs, ptok = ops.xattn_fwd(q, k, vt, B, heads, n, nk, d, scale, loss=kl2, want_probs=True,
                                        save_tok=save_tok, want_lse=True)
    ref3 = ops.linear(o, wo, bo, res)
    torch.cuda.synchronize()
    assert torch.equal(q_f, q)
    assert _rel(out, ref3) < 2e-3, _rel(out, ref3)
    assert (probs_f.float() - probs.float()).abs().max() < 1e-3
    assert (lse_f - lse).abs().max() < 1e-3
    assert (ptok_f.float() - ptok.float()).abs().max() < 1e-3
    if with_loss:
        a, b2 = kl.loss_part.cpu(), kl2.loss_part.cpu()
        assert (a - b2).abs().max() < 2e-4 * max(1.0, float(b2.abs().max())), (a, b2)
        assert _rel(kl.dp_extra, kl2.dp_extra) < 2e-3
    # fp32 torch
    qh = (x.float() @ wq.float().t()).half().float().view(B, n, heads, d).permute(0, 2, 1, 3)
    kh = k[:, :nk, :d].float().view(B, heads, nk, d)
    vh = v[:, :nk, :d].float().view(B, heads, nk, d)
    P = torch.softmax(qh @ kh.transpose(-1, -2) * scale, dim=-1)
    oo = (P @ vh).permute(0, 2, 1, 3).reshape(B * n, C)
    ref = oo @ wo.float().t() + bo + res.float()
    assert _rel(out, ref) < 5e-3, _rel(out, ref)
