"""GPU parity of the tcgen05 attention backward (dQ, dK, dV, and the dP_extra / in-kernel-delta path used by the
guidance loss) against torch autograd in fp32 on the same fp16-rounded operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def _slabs(B, heads, d, n, dev):
    from lgd_b200 import ops
    dp, d16 = ops.round_dp(d), ops.round_d16(d)
    na = (n + 7) // 8 * 8
    rm = lambda: torch.zeros(B * heads, na, dp, device=dev, dtype=torch.float16)
    tr = lambda: torch.zeros(B * heads, d16, na, device=dev, dtype=torch.float16)
    return rm, tr


@pytest.mark.parametrize("B,heads,d,nq,nk,nk_store", [(2, 8, 40, 1024, 1024, None), (1, 8, 80, 256, 256, None),
                                                      (2, 8, 160, 256, 256, None), (1, 8, 160, 64, 64, None),
                                                      (1, 8, 16, 256, 286, 256), (2, 5, 64, 320, 320, None),
                                                      (1, 8, 32, 64, 94, 64),
                                                      # the shapes the guidance step runs at SD1.5 widths, batch 8:
                                                      # 64x64 self-attention and the fuser's 4096+30 ragged tokens
                                                      (8, 8, 40, 4096, 4096, None), (2, 8, 40, 4126, 4126, None),
                                                      (2, 8, 80, 1054, 1054, None), (2, 8, 160, 286, 286, None)])
def test_self_attention_bwd(cuda, B, heads, d, nq, nk, nk_store):
    from lgd_b200 import ops
    C = heads * d
    g = torch.Generator(device="cpu").manual_seed(d * 3 + nq)
    xq = torch.randn(B * nq, C, generator=g).half().to(cuda)
    xk = torch.randn(B * nk, C, generator=g).half().to(cuda)
    w = (torch.randn(3 * C, C, generator=g) * 1.5 / C ** 0.5).half().to(cuda)
    wo = (torch.randn(C, C, generator=g) / C ** 0.5).half().to(cuda)      # used as the dO producer
    dy = torch.randn(B * nq, C, generator=g).half().to(cuda)
    rmq, trq = _slabs(B, heads, d, nq, cuda)
    rmk, trk = _slabs(B, heads, d, nk, cuda)
    q, qt = rmq(), trq()
    k, kt, v, vt = rmk(), trk(), rmk(), trk()
    ops.project_heads2(xq, w[:C], nq, heads, d, 0, rm=(q, None, None), tr=(qt, None, None))
    ops.project_heads2(xk, w[C:], nk, heads, d, 1, rm=(None, k, v), tr=(None, kt, vt))
    scale = d ** -0.5
    out, lse = ops.attention_fwd(q, k, vt, B, heads, nq, nk, d, scale, want_lse=True)
    dO, dOt = rmq(), trq()
    ops.project_heads2(dy, wo, nq, heads, d, 0, rm=(dO, None, None), tr=(dOt, None, None))
    do_tok = ops.linear(dy, wo)
    dq, dk, dv = ops.attention_bwd(q, k, v, dO, qt, kt, dOt, lse, out, do_tok, B, heads, nq, nk, d, scale,
                                   nk_store=nk_store)
    # reference: autograd over the same fp16-rounded q,k,v
    qh = q[:, :nq, :d].float().requires_grad_(True)
    kh = k[:, :nk, :d].float().requires_grad_(True)
    vh = v[:, :nk, :d].float().requires_grad_(True)
    o = torch.softmax(qh @ kh.transpose(1, 2) * scale, dim=-1) @ vh
    o.backward(dO[:, :nq, :d].float())
    unflat = lambda t, n: t.view(B, heads, n, d).permute(0, 2, 1, 3).reshape(B * n, C)
    assert _rel(dq, unflat(qh.grad, nq)) < 8e-3, _rel(dq, unflat(qh.grad, nq))
    ns = nk if nk_store is None else nk_store
    dk_ref = unflat(kh.grad, nk).view(B, nk, C)[:, :ns].reshape(B * ns, C)
    dv_ref = unflat(vh.grad, nk).view(B, nk, C)[:, :ns].reshape(B * ns, C)
    assert _rel(dk, dk_ref) < 8e-3, _rel(dk, dk_ref)
    assert _rel(dv, dv_ref) < 8e-3, _rel(dv, dv_ref)


@pytest.mark.parametrize("B,heads,d,nq,has_do", [(2, 8, 160, 256, True), (2, 8, 160, 256, False), (3, 8, 160, 64, True),
                                                 (1, 8, 64, 256, True), (2, 8, 40, 1024, True), (1, 8, 32, 64, False)])
def test_cross_attention_bwd_with_extra(cuda, B, heads, d, nq, has_do):
    """nk = 77, K/V constant: dQ only, dP = dO V^T + dP_extra, delta formed in-kernel"""
    from lgd_b200 import ops
    C = heads * d
    nk = 77
    g = torch.Generator(device="cpu").manual_seed(d + nq)
    x = torch.randn(B * nq, C, generator=g).half().to(cuda)
    ctx = torch.randn(B * nk, 768, generator=g).half().to(cuda)
    wq = (torch.randn(C, C, generator=g) * 2 / C ** 0.5).half().to(cuda)
    wkv = (torch.randn(2 * C, 768, generator=g) * 1.5 / 768 ** 0.5).half().to(cuda)
    wo = (torch.randn(C, C, generator=g) / C ** 0.5).half().to(cuda)
    dy = torch.randn(B * nq, C, generator=g).half().to(cuda)
    extra = torch.zeros(B * heads, nq, 80, device=cuda)
    extra[:, :, :nk] = torch.randn(B * heads, nq, nk, generator=g).to(cuda) * (torch.rand(B * heads, nq, nk, generator=g).to(cuda) > 0.9)
    rmq, trq = _slabs(B, heads, d, nq, cuda)
    rmk, trk = _slabs(B, heads, d, nk, cuda)
    q = rmq()
    k, kt, v, vt = rmk(), trk(), rmk(), trk()
    ops.project_heads2(x, wq, nq, heads, d, 0, rm=(q, None, None))
    ops.project_heads2(ctx, wkv, nk, heads, d, 1, rm=(None, k, v), tr=(None, kt, vt))
    scale = d ** -0.5
    out, lse = ops.attention_fwd(q, k, vt, B, heads, nq, nk, d, scale, want_lse=True)
    dO = None
    if has_do:
        dO = rmq()
        ops.project_heads2(dy, wo, nq, heads, d, 0, rm=(dO, None, None))
    dq, _, _ = ops.attention_bwd(q, k, v, dO, None, kt, None, lse, None, None, B, heads, nq, nk, d, scale,
                                 dp_extra=extra, want_dkv=False, use_delta=False)
    qh = q[:, :nq, :d].float().requires_grad_(True)
    kh, vh = k[:, :nk, :d].float(), v[:, :nk, :d].float()
    P = torch.softmax(qh @ kh.transpose(1, 2) * scale, dim=-1)
    obj = (P * extra[:, :, :nk]).sum()
    if has_do:
        obj = obj + ((P @ vh) * dO[:, :nq, :d].float()).sum()
    obj.backward()
    ref = qh.grad.view(B, heads, nq, d).permute(0, 2, 1, 3).reshape(B * nq, C)
    assert _rel(dq, ref) < 8e-3, _rel(dq, ref)
