"""GPU parity of the batched denoising loops (llm-groundeddiffusion_b200/pipelines.py) against the oracle loops
(oracle/pipeline_ref.py, pinned to the reference's models/pipelines.py): iteration counts must match exactly (integer
artefacts), loss traces and final latents within the fp16-vs-fp32 tolerance stated per check."""
import pytest
import torch

pytestmark = pytest.mark.gpu
KEYS = [("mid", 0, 0, 0), ("up", 1, 0, 0), ("up", 1, 1, 0), ("up", 1, 2, 0)]


def _rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def _common(gligen, B=2, side=32, seed=0):
    from lgd_b200.unet import B200UNet, UNetConfig
    from oracle import unet_ref
    ocfg = unet_ref.UNetConfig.tiny(gligen=gligen)
    w = unet_ref.make_weights(ocfg, seed=seed)
    net = B200UNet(UNetConfig.tiny(gligen=gligen), w, "cuda:0")
    g = torch.Generator().manual_seed(seed + 5)
    z0 = torch.randn(B, 4, side, side, generator=g)
    uncond = torch.randn(1, 77, 768, generator=g)
    cond = torch.randn(B, 77, 768, generator=g)
    return ocfg, w, net, g, z0, uncond, cond


def test_semantic_guidance_loop(cuda):
    from lgd_b200 import guidance as G, pipelines as P
    from oracle import pipeline_ref
    B, steps = 2, 4
    ocfg, w, net, g, z0, uncond, cond = _common(False, B)
    lay = [G.SampleLayout([[(0.1, 0.2, 0.6, 0.7)], [(0.5, 0.4, 0.95, 0.9)]], [[2, 3], [6]], [3, 6]),
           G.SampleLayout([[(0.3, 0.1, 0.9, 0.5), (0.0, 0.6, 0.4, 1.0)]], [[4]], [4])]
    spec = P.GuidanceSpec(layouts=lay, keys=KEYS, loss_scale=10, loss_threshold=0.2, max_iter=[2, 1, 1],
                          max_index_step=3, fg_weight=1.0, bg_weight=4.0)
    res = P.denoise(net, z0, uncond, cond, steps, guidance=spec, save_keys=[("down", 2, 1, 0)] + KEYS,
                    save_tok=[3, 4], save_latents=True)
    torch.cuda.synchronize()
    iters = list(zip(*res["state"].iters))     # per image
    for b in range(B):
        og = pipeline_ref.GuidanceCfg(lay[b].bboxes, lay[b].object_positions, KEYS, 10, 0.2, [2, 1, 1], 3, 0.2, 0.2,
                                      1.0, 4.0)
        tr = []
        ref = pipeline_ref.denoise(w, ocfg, z0[b:b + 1], uncond, cond[b:b + 1], steps, g=og,
                                   save_keys=[("down", 2, 1, 0)] + KEYS, save_token=[3, 4][b], trace=tr)
        assert list(iters[b]) == ref["iters"], (iters[b], ref["iters"])
        r = _rel(res["latents"][b:b + 1].cpu(), ref["latents"])
        print("image", b, "final-latent rel-L2", r, "oracle loss", ref["loss"], "ours", res["state"].loss[b])
        # fp16 path vs fp32 oracle after several large guidance steps; top-k membership near ties differs between
        # fp16 and fp32 maps (the autocast reference has the same sensitivity), so the bound is loose
        assert r < 0.1, r            # measured 4.8e-2 / 5.0e-2 (large scale-10 guidance steps on the small topology)
        # the first loss evaluation starts from identical latents: tight; the final one has gone through all guidance
        # updates (discrete top-k membership amplifies fp16-vs-fp32 differences): loose
        first_ours, first_ref = res["state"].trace[0][2][b], tr[0][2]
        print("first loss ours", first_ours, "oracle", first_ref)
        assert abs(first_ours - first_ref) < 1e-2 * abs(first_ref)
        assert abs(res["state"].loss[b] - ref["loss"]) < 0.05 * abs(ref["loss"])   # measured 5e-4 / 2e-2
        for si, (s_ref, s) in enumerate(zip(ref["saved"], res["saved"])):
            for k in s_ref:
                # random-weight attention is nearly one-hot, so a near-tie that flips moves single entries by ~1:
                # the saved maps are compared on their mean absolute difference, not the maximum
                d = (s_ref[k][0, :, :, 0] - s[k][b].float().cpu()).abs()
                assert d.mean() < 0.02, (si, k, d.mean(), d.max())


def test_gligen_ref_frozen_loop(cuda):
    from lgd_b200 import guidance as G, pipelines as P
    from oracle import pipeline_ref
    B, steps, side = 2, 4, 32
    ocfg, w, net, g, z0, uncond, cond = _common(True, B)
    frozen = torch.randn(steps + 1, B, 4, side, side, generator=g)
    frozen[0] = z0
    fmask = (torch.rand(B, side, side, generator=g) > 0.5).float()
    boxes = [[(0.1, 0.2, 0.6, 0.7), (0.5, 0.4, 0.95, 0.9)], [(0.2, 0.2, 0.8, 0.8)]]
    lay = [G.SampleLayout([[boxes[0][0]], [boxes[0][1]]], [[2, 3], [6]], [3, 6]),
           G.SampleLayout([[boxes[1][0]]], [[5, 6]], [6])]
    heads = 8
    refs = [[[[{k: torch.softmax(3 * torch.randn(heads, 16 if k[0] == "mid" else 64, generator=g), dim=1)
                for k in KEYS} for _ in range(steps)] for _ in phrase] for phrase in l.bboxes] for l in lay]
    gl = dict(boxes=torch.zeros(B, 30, 4), masks=torch.zeros(B, 30), positive_embeddings=torch.zeros(B, 30, 768))
    for b in range(B):
        n = len(boxes[b])
        gl["boxes"][b, :n] = torch.tensor(boxes[b])
        gl["masks"][b, :n] = 1
        gl["positive_embeddings"][b, :n] = torch.randn(n, 768, generator=g)
    spec = P.GuidanceSpec(layouts=lay, keys=KEYS, loss_scale=5, loss_threshold=0.01, max_iter=[2, 1],
                          max_index_step=3, fg_weight=1.0, bg_weight=4.0, ref_ca_loss_weight=2.0,
                          ref_word_token_only=True, ref_maps=refs)
    res = P.denoise(net, z0, uncond, cond, steps, guidance=spec, frozen_mask=fmask, frozen_latents=frozen,
                    frozen_steps=2, gligen=gl, gligen_beta=0.5)
    torch.cuda.synchronize()
    iters = list(zip(*res["state"].iters))
    for b in range(B):
        og = pipeline_ref.GuidanceCfg(lay[b].bboxes, lay[b].object_positions, KEYS, 5, 0.01, [2, 1], 3, 0.2, 0.2, 1.0,
                                      4.0, refs[b], lay[b].word_token_indices, 2.0, True)
        ref = pipeline_ref.denoise(w, ocfg, z0[b:b + 1], uncond, cond[b:b + 1], steps, g=og, frozen_mask=fmask[b],
                                   frozen_latents=frozen[:, b:b + 1], frozen_steps=2,
                                   gligen={k: v[b:b + 1] for k, v in gl.items()}, gligen_beta=0.5)
        assert list(iters[b]) == ref["iters"], (iters[b], ref["iters"])
        r = _rel(res["latents"][b:b + 1].cpu(), ref["latents"])
        print("image", b, "final-latent rel-L2", r, "oracle loss", ref["loss"], "ours", res["state"].loss[b])
        # fp16 path vs fp32 oracle after several large guidance steps; top-k membership near ties differs between
        # fp16 and fp32 maps (the autocast reference has the same sensitivity), so the bound is loose
        assert r < 0.15, r           # measured 3.5e-2 / 7.8e-2
        assert abs(res["state"].loss[b] - ref["loss"]) < 0.05 * abs(ref["loss"])   # measured 5e-4 / 6e-3


def test_fast_schedule_loop(cuda):
    """thinned timestep list + per-step DDIM step size (use_fast_schedule of generation/lmd_plus.py:360-367,
    models/pipelines.py:358-362,439-440,449) vs the oracle loop"""
    from lgd_b200 import pipelines as P
    from oracle import pipeline_ref
    B, steps = 2, 8
    ocfg, w, net, g, z0, uncond, cond = _common(True, B)
    gl = dict(boxes=torch.zeros(B, 30, 4), masks=torch.zeros(B, 30), positive_embeddings=torch.zeros(B, 30, 768))
    for b in range(B):
        gl["boxes"][b, 0] = torch.tensor([0.1 + 0.2 * b, 0.2, 0.6 + 0.2 * b, 0.7])
        gl["masks"][b, 0] = 1
        gl["positive_embeddings"][b, 0] = torch.randn(768, generator=g)
    res = P.denoise(net, z0, uncond, cond, steps, gligen=gl, gligen_beta=0.5, save_latents=True, fast_after_steps=3,
                    fast_rate=2, dynamic_num_inference_steps=True)
    torch.cuda.synchronize()
    assert res["latents_all"].shape[0] == 4
    for b in range(B):
        ref = pipeline_ref.denoise(w, ocfg, z0[b:b + 1], uncond, cond[b:b + 1], steps,
                                   gligen={k: v[b:b + 1] for k, v in gl.items()}, gligen_beta=0.5, fast_after_steps=3,
                                   fast_rate=2, dynamic_num_inference_steps=True)
        r = _rel(res["latents"][b:b + 1].cpu(), ref["latents"])
        print("image", b, "final-latent rel-L2", r)
        assert r < 1e-2, r           # measured 4.4e-3 / 4.5e-3
        assert _rel(res["latents_all"][:, b:b + 1].cpu(), ref["latents_all"]) < 2e-2


def test_v_prediction_guidance_loop(cuda):
    """BASELINE config 3 semantics on a small topology: SD2.1-style UNet (head_dim 64, linear projections, 1024-wide
    context), v-prediction DDIM step, backward-guidance constants (loss_scale 30, threshold 0.2, max_iter 5,
    max_index_step 10 -> here 2)"""
    from lgd_b200 import guidance as G, pipelines as P
    from lgd_b200.unet import B200UNet, UNetConfig
    from oracle import pipeline_ref, unet_ref
    kw = dict(block_out_channels=(128, 256, 512, 512), heads=(2, 4, 8, 8), cross_attention_dim=1024,
              use_linear_projection=True)
    ocfg = unet_ref.UNetConfig(**kw)
    w = unet_ref.make_weights(ocfg, seed=2)
    net = B200UNet(UNetConfig(**kw), w, "cuda:0")
    B, steps, side = 2, 3, 32
    g = torch.Generator().manual_seed(8)
    z0 = torch.randn(B, 4, side, side, generator=g)
    uncond = torch.randn(1, 77, 1024, generator=g)
    cond = torch.randn(B, 77, 1024, generator=g)
    lay = [G.SampleLayout([[(0.1, 0.2, 0.6, 0.7)]], [[2, 3]], [3]),
           G.SampleLayout([[(0.4, 0.3, 0.9, 0.8)]], [[5]], [5])]
    spec = P.GuidanceSpec(layouts=lay, keys=KEYS, loss_scale=3, loss_threshold=0.2, max_iter=2, max_index_step=2,
                          fg_weight=1.0, bg_weight=1.0)
    res = P.denoise(net, z0, uncond, cond, steps, guidance=spec, prediction_type="v_prediction")
    torch.cuda.synchronize()
    iters = list(zip(*res["state"].iters))
    for b in range(B):
        og = pipeline_ref.GuidanceCfg(lay[b].bboxes, lay[b].object_positions, KEYS, 3, 0.2, 2, 2, 0.2, 0.2, 1.0, 1.0)
        ref = pipeline_ref.denoise(w, ocfg, z0[b:b + 1], uncond, cond[b:b + 1], steps, g=og,
                                   prediction_type="v_prediction")
        assert list(iters[b]) == ref["iters"], (iters[b], ref["iters"])
        r = _rel(res["latents"][b:b + 1].cpu(), ref["latents"])
        print("image", b, "final-latent rel-L2", r, "oracle loss", ref["loss"], "ours", res["state"].loss[b])
        assert r < 0.04, r           # measured 1.6e-2 / 1.7e-2
        assert abs(res["state"].loss[b] - ref["loss"]) < 0.02 * abs(ref["loss"])   # measured 7e-4 / 2.5e-3
