"""GPU parity of the fused cross-attention + guidance-loss kernel against the oracle (oracle/guidance_ref.py, itself
pinned to the reference's utils/guidance.py) on the kernel's own fp16 attention maps, and of the loss gradient
(dP_extra -> dQ through attn_bwd_dq) against autograd through the oracle loss."""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
KEY = ("up", 1, 0, 0)


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def _layouts(B, seed, heads, n, with_ref):
    from lgd_b200 import guidance
    rng = random.Random(seed)
    g = torch.Generator().manual_seed(seed)
    out = []
    for b in range(B):
        n_obj = rng.randint(1, 4)
        bboxes, pos, words, refs = [], [], [], []
        tok = 1
        for o in range(n_obj):
            nb = rng.randint(1, 2)
            boxes = []
            for _ in range(nb):
                w_, h_ = rng.uniform(0.1, 0.6), rng.uniform(0.1, 0.6)
                x, y = rng.uniform(0, 1 - w_), rng.uniform(0, 1 - h_)
                boxes.append((x, y, x + w_, y + h_))
            bboxes.append(boxes)
            nt = rng.randint(1, 3)
            pos.append(list(range(tok, tok + nt)))
            words.append(tok + nt - 1)
            tok += nt + rng.randint(0, 1)     # sometimes adjacent phrases share no gap
            refs.append([{KEY: torch.softmax(3 * torch.randn(heads, n, generator=g), dim=1).numpy()}
                         for _ in range(nb)])
        out.append(guidance.SampleLayout(bboxes, pos, words, refs if with_ref else None))
    return out


@pytest.mark.parametrize("B,heads,d,n,with_ref", [(2, 8, 160, 256, False), (3, 8, 160, 256, True), (4, 8, 160, 64, True),
                                                  (2, 8, 32, 256, True), (1, 8, 64, 1024, False)])
def test_fused_xattn_loss_and_grad(cuda, B, heads, d, n, with_ref):
    from lgd_b200 import guidance, ops
    from oracle import guidance_ref
    C, nk = heads * d, 77
    g = torch.Generator(device="cpu").manual_seed(n + d)
    x = torch.randn(B * n, C, generator=g).half().to(cuda)
    ctx = torch.randn(B * nk, 768, generator=g).half().to(cuda)
    wq = (torch.randn(C, C, generator=g) * 3 / C ** 0.5).half().to(cuda)
    wkv = (torch.randn(2 * C, 768, generator=g) * 2 / 768 ** 0.5).half().to(cuda)
    dp, d16 = ops.round_dp(d), ops.round_d16(d)
    z = lambda *s: torch.zeros(*s, device=cuda, dtype=torch.float16)
    q = z(B * heads, n, dp)
    k, v, kt, vt = z(B * heads, 80, dp), z(B * heads, 80, dp), z(B * heads, d16, 80), z(B * heads, d16, 80)
    ops.project_heads2(x, wq, n, heads, d, 0, rm=(q, None, None))
    ops.project_heads2(ctx, wkv, nk, heads, d, 1, rm=(None, k, v), tr=(None, kt, vt))
    samples = _layouts(B, n + d, heads, n, with_ref)
    params = guidance.LossParams(loss_scale=5.0, fg_top_p=0.2, bg_top_p=0.2, fg_weight=1.0, bg_weight=4.0,
                                 ref_ca_loss_weight=2.0, ref_word_token_only=True, use_ref=with_ref)
    slot_tok, slot_of = guidance.assign_slots(samples, params)
    slot_dev = torch.from_numpy(slot_tok).to(cuda)
    n_keys = 4
    gscale = 64.0
    kl = guidance.KeyLoss(samples, slot_dev, slot_of, KEY, n, heads, n_keys, params, cuda, gscale=gscale)
    scale = d ** -0.5
    save_tok = torch.tensor([s.object_positions[0][-1] for s in samples], dtype=torch.int32, device=cuda)
    out, lse, probs, ptok = ops.xattn_fwd(q, k, vt, B, heads, n, nk, d, scale, loss=kl, want_probs=True,
                                          save_tok=save_tok, want_lse=True)
    torch.cuda.synchronize()
    # attention output and maps
    qh, kh, vh = q[:, :, :d].float(), k[:, :nk, :d].float(), v[:, :nk, :d].float()
    P = torch.softmax(qh @ kh.transpose(1, 2) * scale, dim=-1)
    ref_out = (P @ vh).view(B, heads, n, d).permute(0, 2, 1, 3).reshape(B * n, C)
    assert _rel(out, ref_out) < 4e-3
    assert (probs.float() - P).abs().max() < 2e-3
    for b in range(B):
        col = probs.view(B, heads, n, nk)[b, :, :, int(save_tok[b])]
        assert torch.equal(ptok.view(B, heads, n)[b], col)
    # loss on the kernel's own fp16 maps vs the oracle (the oracle's normalisers include 1/n_keys for 1 key => x1/n_keys)
    loss_dev = kl.loss_per_image().cpu()
    probs_cpu = probs.float().cpu().view(B, heads, n, nk)
    for b, s in enumerate(samples):
        refs = None
        if with_ref:
            refs = [[{KEY: torch.from_numpy(m[KEY])} for m in obj] for obj in s.ref_maps]
        L = guidance_ref.ca_loss({KEY: probs_cpu[b]}, s.bboxes, s.object_positions, [KEY], 0.2, 0.2, 1.0, 4.0, refs,
                                 s.word_token_indices, 2.0, True)
        expect = float(L) * 5.0 / n_keys
        assert abs(float(loss_dev[b]) - expect) < 2e-4 * max(1.0, abs(expect)), (b, float(loss_dev[b]), expect)
        L2, grads = guidance_ref.ca_loss_and_grad(
            {KEY: probs_cpu[b].numpy()}, s.bboxes, s.object_positions, [KEY], 0.2, 0.2, 1.0, 4.0,
            None if refs is None else [[{KEY: m[KEY].numpy()} for m in obj] for obj in refs], s.word_token_indices,
            2.0, True)
        dpx = kl.dp_extra.view(B, heads, n, 80)[b, :, :, :nk].cpu().numpy() / gscale
        np.testing.assert_allclose(dpx, grads[KEY] * 5.0 / n_keys, rtol=2e-3, atol=2e-7)
    # d loss / d Q through the backward kernel vs autograd through the oracle loss
    dq, _, _ = ops.attention_bwd(q, k, v, None, None, kt, None, lse, None, None, B, heads, n, nk, d, scale,
                                 dp_extra=kl.dp_extra, want_dkv=False, use_delta=False)
    qg = q[:, :, :d].float().cpu().requires_grad_(True)
    Pg = torch.softmax(qg @ kh.cpu().transpose(1, 2) * scale, dim=-1).view(B, heads, n, nk)
    tot = 0
    for b, s in enumerate(samples):
        refs = None
        if with_ref:
            refs = [[{KEY: torch.from_numpy(m[KEY])} for m in obj] for obj in s.ref_maps]
        tot = tot + guidance_ref.ca_loss({KEY: Pg[b]}, s.bboxes, s.object_positions, [KEY], 0.2, 0.2, 1.0, 4.0, refs,
                                         s.word_token_indices, 2.0, True) * 5.0 / n_keys
    tot.backward()
    ref_dq = qg.grad.view(B, heads, n, d).permute(0, 2, 1, 3).reshape(B * n, C) * gscale
    assert _rel(dq.cpu(), ref_dq) < 2e-2, _rel(dq.cpu(), ref_dq)


@pytest.mark.parametrize("B,heads,d,n", [(2, 8, 160, 256), (3, 8, 160, 64), (2, 5, 64, 256), (1, 8, 40, 1024)])
def test_ratio_based_loss_and_grad(cuda, B, heads, d, n):
    """loss term type 2 - the ratio-based energy of utils/guidance.py:122-128 that backward_guidance.py runs - in the
    cross-attention kernel: per-image loss and the dense d loss / d P vs autograd through the oracle"""
    from lgd_b200 import guidance, ops
    from oracle import guidance_ref
    C, nk = heads * d, 77
    g = torch.Generator(device="cpu").manual_seed(n + d + 1)
    x = torch.randn(B * n, C, generator=g).half().to(cuda)
    ctx = torch.randn(B * nk, 768, generator=g).half().to(cuda)
    wq = (torch.randn(C, C, generator=g) * 3 / C ** 0.5).half().to(cuda)
    wkv = (torch.randn(2 * C, 768, generator=g) * 2 / 768 ** 0.5).half().to(cuda)
    dp, d16 = ops.round_dp(d), ops.round_d16(d)
    z = lambda *s: torch.zeros(*s, device=cuda, dtype=torch.float16)
    q = z(B * heads, n, dp)
    k, v, kt, vt = z(B * heads, 80, dp), z(B * heads, 80, dp), z(B * heads, d16, 80), z(B * heads, d16, 80)
    ops.project_heads2(x, wq, n, heads, d, 0, rm=(q, None, None))
    ops.project_heads2(ctx, wkv, nk, heads, d, 1, rm=(None, k, v), tr=(None, kt, vt))
    samples = _layouts(B, n + d + 1, heads, n, False)
    params = guidance.LossParams(loss_scale=30.0, use_ratio_based_loss=True)
    slot_tok, slot_of = guidance.assign_slots(samples, params)
    n_keys, gscale = 4, 64.0
    kl = guidance.KeyLoss(samples, torch.from_numpy(slot_tok).to(cuda), slot_of, KEY, n, heads, n_keys, params, cuda,
                          gscale=gscale)
    out, lse, probs, _ = ops.xattn_fwd(q, k, vt, B, heads, n, nk, d, d ** -0.5, loss=kl, want_probs=True, want_lse=True)
    torch.cuda.synchronize()
    loss_dev = kl.loss_per_image().cpu()
    probs_cpu = probs.float().cpu().view(B, heads, n, nk)
    for b, s in enumerate(samples):
        Pb = probs_cpu[b].clone().requires_grad_(True)
        L = guidance_ref.ca_loss({KEY: Pb}, s.bboxes, s.object_positions, [KEY], use_ratio_based_loss=True) * 30.0 / n_keys
        gP = torch.autograd.grad(L, [Pb])[0]
        assert abs(float(loss_dev[b]) - float(L)) < 2e-4 * max(1.0, abs(float(L))), (b, float(loss_dev[b]), float(L))
        dpx = kl.dp_extra.view(B, heads, n, 80)[b, :, :, :nk].cpu() / gscale
        assert _rel(dpx, gP) < 2e-3, _rel(dpx, gP)
