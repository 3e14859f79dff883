"""GPU parity: head-split projection + tcgen05 attention forward, GroupNorm/LayerNorm/GEGLU forward+backward, against
plain PyTorch fp32 on the same fp16-rounded inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


@pytest.mark.parametrize("B,heads,d,nq,nk,cross", [
    (2, 8, 40, 4096, 4096, False), (2, 8, 80, 1024, 1024, False), (2, 8, 160, 256, 256, False),
    (3, 8, 160, 64, 64, False), (2, 8, 40, 4096, 77, True), (2, 8, 160, 256, 77, True), (1, 8, 16, 256, 286, False),
    (2, 5, 64, 1024, 1024, False), (1, 8, 32, 64, 94, False), (2, 8, 64, 576, 576, False),
    (2, 8, 40, 4126, 4126, False), (2, 8, 80, 1054, 1054, False), (2, 8, 160, 286, 286, False),   # fuser: n + 30
    (1, 5, 64, 9216, 9216, False)])                                                                  # SD2.1 at 96x96
def test_attention_fwd(cuda, B, heads, d, nq, nk, cross):
    from lgd_b200 import ops
    C = heads * d
    g = torch.Generator(device="cpu").manual_seed(d + nq)
    kin = 768 if cross else C
    x = torch.randn(B * nq, C, generator=g).half().to(cuda)
    src = torch.randn(B * nk, kin, generator=g).half().to(cuda) if (cross or nk != nq) else x
    wq = (torch.randn(C, C, generator=g) * 2 / C ** 0.5).half().to(cuda)
    wk = (torch.randn(C, src.shape[1], generator=g) * 2 / src.shape[1] ** 0.5).half().to(cuda)
    wv = (torch.randn(C, src.shape[1], generator=g) / src.shape[1] ** 0.5).half().to(cuda)
    nk_alloc = (nk + 7) // 8 * 8
    q, k, vt = ops.alloc_head_slabs(B, heads, d, nq, nk_alloc, cuda)
    ops.project_heads(x, wq, nq, heads, d, 0, q=q)
    ops.project_heads(src, torch.cat([wk, wv], 0), nk, heads, d, 1, k=k, vt=vt)
    # check the slabs themselves
    qr = (x.float() @ wq.float().t()).view(B, nq, heads, d).permute(0, 2, 1, 3).reshape(B * heads, nq, d)
    kr = (src.float() @ wk.float().t()).view(B, nk, heads, d).permute(0, 2, 1, 3).reshape(B * heads, nk, d)
    vr = (src.float() @ wv.float().t()).view(B, nk, heads, d).permute(0, 2, 1, 3).reshape(B * heads, nk, d)
    assert _rel(q[:, :, :d], qr) < 2e-3
    assert _rel(k[:, :nk, :d], kr) < 2e-3
    assert _rel(vt[:, :d, :nk], vr.transpose(1, 2)) < 2e-3
    assert float(q[:, :, d:].abs().max() if q.shape[2] > d else 0) == 0
    scale = d ** -0.5
    out, lse = ops.attention_fwd(q, k, vt, B, heads, nq, nk, d, scale, want_lse=True)
    qh, kh, vh = q[:, :, :d].float(), k[:, :nk, :d].float(), vt[:, :d, :nk].float().transpose(1, 2)
    s = qh @ kh.transpose(1, 2) * scale
    pr = torch.softmax(s, dim=-1)
    ref = (pr @ vh).view(B, heads, nq, d).permute(0, 2, 1, 3).reshape(B * nq, C)
    assert _rel(out, ref) < 4e-3, _rel(out, ref)
    lse_ref = torch.logsumexp(s, dim=-1) * 1.4426950408889634
    assert (lse[:, :nq] - lse_ref).abs().max() < 2e-3


@pytest.mark.parametrize("B,n,C,silu", [(2, 4096, 320, True), (2, 256, 1280, True), (3, 64, 2560, False),
                                        (1, 1024, 1920, True), (2, 256, 128, True), (2, 1024, 960, False)])
def test_groupnorm_fwd_bwd(cuda, B, n, C, silu):
    from lgd_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(C)
    x = (torch.randn(B, n, C, generator=g) * 1.5 + 0.3).half().to(cuda)
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).to(cuda)
    beta = (0.2 * torch.randn(C, generator=g)).to(cuda)
    dy = torch.randn(B, n, C, generator=g).half().to(cuda)
    eps = 1e-5
    y, sums = ops.groupnorm(x, gamma, beta, 32, eps, silu, want_sums=True)
    xr = x.float().requires_grad_(True)
    t = torch.nn.functional.group_norm(xr.transpose(1, 2), 32, gamma, beta, eps).transpose(1, 2)
    ref = torch.nn.functional.silu(t) if silu else t
    assert _rel(y, ref) < 2e-3
    ref.backward(dy.float())
    dx = ops.groupnorm_bwd(dy, x, sums, gamma, beta, 32, eps, silu)
    assert _rel(dx, xr.grad) < 4e-3, _rel(dx, xr.grad)
    dx2 = ops.groupnorm_bwd(dy, x, sums, gamma, beta, 32, eps, silu, dx=dx.clone())
    assert _rel(dx2, 2 * xr.grad) < 4e-3


@pytest.mark.parametrize("rows,C", [(8192, 320), (2048, 640), (512, 1280), (100, 128)])
def test_layernorm_fwd_bwd(cuda, rows, C):
    from lgd_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(C)
    x = (torch.randn(rows, C, generator=g) * 2 + 0.5).half().to(cuda)
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).to(cuda)
    beta = (0.2 * torch.randn(C, generator=g)).to(cuda)
    dy = torch.randn(rows, C, generator=g).half().to(cuda)
    y, stats = ops.layernorm(x, gamma, beta, want_stats=True)
    xr = x.float().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (C,), gamma, beta, 1e-5)
    assert _rel(y, ref) < 2e-3
    ref.backward(dy.float())
    dx = ops.layernorm_bwd(dy, x, stats, gamma)
    assert _rel(dx, xr.grad) < 3e-3


def test_geglu_bwd(cuda):
    from lgd_b200 import ops
    M, F, K = 256, 1280, 320
    g = torch.Generator(device="cpu").manual_seed(3)
    x = torch.randn(M, K, generator=g).half().to(cuda)
    w = (torch.randn(2 * F, K, generator=g) / K ** 0.5).half().to(cuda)
    b = torch.randn(2 * F, generator=g).to(cuda)
    dy = torch.randn(M, F, generator=g).half().to(cuda)
    w_il, b_il = ops.geglu_interleave(w, b)
    y, pre = ops.linear_geglu(x, w_il, b_il, want_pre=True)
    dpre = ops.geglu_bwd(pre, dy)
    h = (x.float() @ w.float().t() + b).requires_grad_(True)
    v, gate = h.chunk(2, dim=-1)
    (v * torch.nn.functional.gelu(gate)).backward(dy.float())
    dv, dg = h.grad.chunk(2, dim=-1)
    dpre = dpre.float().view(M, F // 64, 2, 64)
    assert _rel(dpre[:, :, 0].reshape(M, F), dv) < 3e-3
    assert _rel(dpre[:, :, 1].reshape(M, F), dg) < 3e-3
