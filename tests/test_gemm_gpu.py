"""GPU parity of the tcgen05 implicit-GEMM family against plain PyTorch fp32 on the same fp16-rounded inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 320, 320), (4096, 320, 320), (300, 200, 136),
                                   (64, 1280, 1280), (2048, 1280, 1280), (77, 640, 768), (1024, 2560, 640),
                                   (4096, 960, 320), (4096, 1920, 640)])   # 256-wide tiles, ragged last tile
def test_linear(cuda, M, N, K):
    from lgd_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N)
    x = (torch.randn(M, K, generator=g) * 0.5).half().to(cuda)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).half().to(cuda)
    bias = torch.randn(N, generator=g).to(cuda)
    res = torch.randn(M, N, generator=g).half().to(cuda)
    y = ops.linear(x, w, bias, res)
    ref = x.float() @ w.float().t() + bias + res.float()
    assert _rel(y, ref) < 2e-3, _rel(y, ref)
    y2, y32 = ops.linear(x, w, None, None, alpha=0.5, out_f32=True)
    ref2 = 0.5 * (x.float() @ w.float().t())
    assert _rel(y32, ref2) < 1e-5
    # accumulate
    y3 = ops.linear(x, w, None, None, out=y2.clone(), accumulate=True)
    assert _rel(y3, ref2 + ref2 * 2) < 3e-3


@pytest.mark.parametrize("M,F,K", [(256, 1280, 320), (64, 5120, 1280), (200, 128, 64)])
def test_geglu(cuda, M, F, K):
    from lgd_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(F)
    x = torch.randn(M, K, generator=g).half().to(cuda)
    w = (torch.randn(2 * F, K, generator=g) / K ** 0.5).half().to(cuda)
    b = torch.randn(2 * F, generator=g).to(cuda)
    w_il, b_il = ops.geglu_interleave(w, b)
    y, pre = ops.linear_geglu(x, w_il, b_il, want_pre=True)
    h = x.float() @ w.float().t() + b
    v, gate = h.chunk(2, dim=-1)
    ref = v * torch.nn.functional.gelu(gate)
    assert _rel(y, ref) < 2e-3, _rel(y, ref)
    # pre-activation dump is tile-interleaved: [t*128 + j] = value t*64+j, [t*128+64+j] = gate
    pre = pre.float().view(M, F // 64, 2, 64)
    assert _rel(pre[:, :, 0].reshape(M, F), v) < 2e-3
    assert _rel(pre[:, :, 1].reshape(M, F), gate) < 2e-3


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(1, 64, 64, 320, 320), (2, 32, 32, 640, 640), (2, 16, 16, 1280, 1280),
                                            (3, 8, 8, 2560, 1280), (1, 8, 8, 1280, 1280), (2, 16, 16, 128, 64),
                                            (1, 24, 24, 192, 320), (2, 64, 64, 320, 4)])
def test_conv3x3(cuda, B, H, W, Cin, Cout):
    from lgd_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(Cin + Cout)
    x = torch.randn(B, H, W, Cin, generator=g).half().to(cuda)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (9 * Cin) ** 0.5).half().to(cuda)
    bias = torch.randn(Cout, generator=g).to(cuda)
    temb = torch.randn(B, Cout, generator=g).to(cuda)
    res = torch.randn(B, H, W, Cout, generator=g).half().to(cuda)
    wk = w.permute(0, 2, 3, 1).reshape(Cout, 9, Cin).contiguous()
    y, y32 = ops.conv3x3(x, wk, bias, temb, res, out_f32=True)
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1)
    ref = ref + temb[:, :, None, None]
    ref = ref.permute(0, 2, 3, 1) + res.float()
    assert _rel(y32, ref) < 1e-4, _rel(y32, ref)
    assert _rel(y, ref) < 2e-3
