"""Pins the oracle restatement (oracle/*.py) against the UNMODIFIED reference imported from /root/reference.

CPU only; runs in the build container (skipped where /root/reference is absent, e.g. on the GPU box).  The reference
has no tests or golden vectors of its own, so outputs of the reference itself, run here on seeded
synthetic weights, are the pin; oracle/make_goldens.py freezes the same cases into tests/golden/.
"""
import random
import sys

import numpy as np
import pytest
import torch

from oracle import guidance_ref, pipeline_ref, ref_loader, unet_ref

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")

KEYS = [("mid", 0, 0, 0), ("up", 1, 0, 0), ("up", 1, 1, 0), ("up", 1, 2, 0)]


def test_scale_proportion_matches_reference():
    r = ref_loader.load()
    rng = random.Random(0)
    for _ in range(3000):
        x0, y0 = rng.uniform(-0.1, 0.9), rng.uniform(-0.1, 0.9)
        box = (x0, y0, x0 + rng.uniform(0, 0.8), y0 + rng.uniform(0, 0.8))
        for side in (8, 16, 24, 64):
            assert guidance_ref.scale_proportion(box, side, side) == r.utils.scale_proportion(box, side, side)
    # .5 cases (banker's rounding)
    for box in [(0.0625, 0.1875, 0.5625, 0.8125), (0.03125, 0.09375, 0.15625, 0.21875)]:
        for side in (8, 16):
            assert guidance_ref.scale_proportion(box, side, side) == r.utils.scale_proportion(box, side, side)


def _random_case(seed, heads=8, with_ref=True):
    g = torch.Generator().manual_seed(seed)
    rng = random.Random(seed)
    saved = {}
    for k in KEYS:
        n = 64 if k[0] == "mid" else 256
        saved[k] = torch.softmax(3 * torch.randn(1, heads, n, 77, generator=g), dim=-1)
    n_obj = rng.randint(1, 3)
    bboxes, positions, words, refs = [], [], [], []
    tok = 1
    for o in range(n_obj):
        nb = rng.randint(1, 2)
        boxes = []
        for _ in range(nb):
            w_, h_ = rng.uniform(0.15, 0.6), rng.uniform(0.15, 0.6)
            x, y = rng.uniform(0, 1 - w_), rng.uniform(0, 1 - h_)
            boxes.append((x, y, x + w_, y + h_))
        bboxes.append(boxes)
        nt = rng.randint(1, 3)
        positions.append(list(range(tok, tok + nt)))
        words.append(tok + nt - 1)
        tok += nt + 1
        refs.append([[{k: torch.softmax(3 * torch.randn(1, heads, saved[k].shape[2], 1, generator=g), dim=2)
                       for k in KEYS}] for _ in range(nb)])   # [box][step=0][key]
    return saved, bboxes, positions, words, (refs if with_ref else None)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
@pytest.mark.parametrize("with_ref", [False, True])
def test_ca_loss_and_grad_match_reference(seed, with_ref):
    r = ref_loader.load()
    saved, bboxes, positions, words, refs = _random_case(seed, with_ref=with_ref)
    leaf = {k: v.clone().requires_grad_(True) for k, v in saved.items()}
    kw = dict(fg_top_p=0.2, bg_top_p=0.2, fg_weight=1.0, bg_weight=4.0, use_ratio_based_loss=False)
    L_ref = r.guidance.compute_ca_lossv3(leaf, bboxes, positions, KEYS, ref_ca_saved_attns=refs, index=0,
                                         ref_ca_word_token_only=True, word_token_indices=words,
                                         ref_ca_loss_weight=2.0, **kw)
    g_ref = torch.autograd.grad(L_ref, [leaf[k] for k in KEYS])
    one = {k: v[0] for k, v in saved.items()}
    ref_maps = None
    if refs is not None:
        ref_maps = [[{k: box[0][k][0, :, :, 0] for k in KEYS} for box in obj] for obj in refs]
    L = guidance_ref.ca_loss(one, bboxes, positions, KEYS, 0.2, 0.2, 1.0, 4.0, ref_maps, words, 2.0, True)
    assert abs(float(L) - float(L_ref)) < 2e-6 * max(1.0, abs(float(L_ref)))
    L2, grads = guidance_ref.ca_loss_and_grad({k: v.numpy() for k, v in one.items()}, bboxes, positions, KEYS, 0.2,
                                              0.2, 1.0, 4.0,
                                              None if ref_maps is None else
                                              [[{k: m[k].numpy() for k in KEYS} for m in obj] for obj in ref_maps],
                                              words, 2.0, True)
    assert abs(L2 - float(L_ref)) < 5e-6 * max(1.0, abs(float(L_ref)))
    for k, gr in zip(KEYS, g_ref):
        np.testing.assert_allclose(grads[k], gr[0].numpy(), rtol=2e-4, atol=2e-7)


@pytest.mark.parametrize("gligen", [False, True])
def test_unet_forward_matches_reference(gligen):
    from oracle import refrun
    cfg = unet_ref.UNetConfig.tiny(gligen=gligen)
    w = unet_ref.make_weights(cfg, seed=0)
    m = refrun.build_reference_unet(cfg, w)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 4, 16, 16, generator=g)
    ctx = torch.randn(2, 77, 768, generator=g)
    kw = {"save_attn_to_dict": {}, "save_keys": None, "enable_flash_attn": False}
    gl = None
    if gligen:
        gl = dict(boxes=torch.rand(2, 30, 4, generator=g), masks=(torch.rand(2, 30, generator=g) > 0.8).float(),
                  positive_embeddings=torch.randn(2, 30, 768, generator=g))
        kw["gligen"] = dict(gl)
    with torch.no_grad():
        ref = m(x, 481, encoder_hidden_states=ctx, cross_attention_kwargs=kw).sample
        saved = {}
        mine = unet_ref.unet_forward(w, cfg, x, 481, ctx, gligen=gl, saved=saved)
    assert (ref - mine).abs().max() < 2e-5
    assert len(saved) == 16
    for k, v in kw["save_attn_to_dict"].items():
        assert (v - saved[k]).abs().max() < 1e-4


def _setup_pipeline(gligen):
    from oracle import refrun
    cfg = unet_ref.UNetConfig.tiny(gligen=gligen)
    w = unet_ref.make_weights(cfg, seed=0)
    g = torch.Generator().manual_seed(5)
    z0 = torch.randn(1, 4, 32, 32, generator=g)
    uncond = torch.randn(1, 77, 768, generator=g)
    cond = torch.randn(1, 77, 768, generator=g)
    table = torch.randn(4, 768, generator=g)
    tok = refrun.FakeTokenizer({"a cat": 0, "a dog": 1})
    enc = refrun.FakeTextEncoder(table)
    r, md = refrun.model_dict(cfg, w, tok, enc)
    return cfg, w, r, md, z0, uncond, cond, table


def test_semantic_guidance_loop_matches_reference():
    """generate_semantic_guidance (LMD per-box phase, backward_guidance): guidance + CFG + DDIM, attention saving"""
    cfg, w, r, md, z0, uncond, cond, _ = _setup_pipeline(False)
    bboxes = [[(0.1, 0.2, 0.6, 0.7)], [(0.5, 0.4, 0.95, 0.9)]]
    positions = [[2, 3], [6]]
    steps = 4
    kw = dict(loss_scale=30, loss_threshold=0.2, max_iter=[2, 1, 1], max_index_step=3, guidance_attn_keys=KEYS,
              fg_top_p=0.2, bg_top_p=0.2, fg_weight=1.0, bg_weight=4.0, use_ratio_based_loss=False,
              ref_ca_loss_weight=0.5, verbose=False)
    out = r.pipelines.generate_semantic_guidance(
        md, z0, (torch.cat([uncond, cond]), uncond, cond), steps, bboxes, ["a cat", "a dog"], positions,
        semantic_guidance_kwargs=kw, return_saved_cross_attn=True, saved_cross_attn_keys=[("down", 2, 1, 0)] + KEYS,
        return_cond_ca_only=True, return_token_ca_only=3, save_all_latents=True, show_progress=False)
    lat_ref, _, saved_ref, all_ref = out
    g = pipeline_ref.GuidanceCfg(bboxes, positions, KEYS, 30, 0.2, [2, 1, 1], 3, 0.2, 0.2, 1.0, 4.0)
    res = pipeline_ref.denoise(w, cfg, z0, uncond, cond, steps, g=g, save_keys=[("down", 2, 1, 0)] + KEYS,
                               save_token=3)
    assert res["iters"] == [2, 1, 1, 0]
    assert (res["latents"] - lat_ref).abs().max() < 5e-3
    assert (res["latents_all"] - all_ref).abs().max() < 5e-3
    for s_ref, s in zip(saved_ref, res["saved"]):
        for k in s_ref:
            assert s_ref[k].shape == s[k].shape
            assert (s_ref[k] - s[k]).abs().max() < 1e-3


def test_partial_frozen_loop_matches_reference():
    """generate_partial_frozen (LMD overall phase, models/pipelines.py:541-599): guidance + CFG + DDIM with the frozen
    blend z = z_ref[i+1] m + z (1 - m) for index < frozen_steps"""
    cfg, w, r, md, z0, uncond, cond, _ = _setup_pipeline(False)
    steps = 4
    g0 = torch.Generator().manual_seed(21)
    latents_all = torch.randn(steps + 1, 1, 4, 32, 32, generator=g0)
    latents_all[0] = z0
    frozen_mask = (torch.rand(32, 32, generator=g0) > 0.4).float()
    bboxes = [[(0.1, 0.2, 0.6, 0.7)], [(0.5, 0.4, 0.95, 0.9)]]
    positions = [[2, 3], [6]]
    kw = dict(loss_scale=30, loss_threshold=0.2, max_iter=[2, 1], max_index_step=3, guidance_attn_keys=KEYS,
              fg_top_p=0.2, bg_top_p=0.2, fg_weight=1.0, bg_weight=4.0, use_ratio_based_loss=False, verbose=False)
    lat_ref, _ = r.pipelines.generate_partial_frozen(
        md, latents_all, frozen_mask, (torch.cat([uncond, cond]), uncond, cond), steps, 2, bboxes=bboxes,
        phrases=["a cat", "a dog"], object_positions=positions, semantic_guidance_kwargs=kw)
    g = pipeline_ref.GuidanceCfg(bboxes, positions, KEYS, 30, 0.2, [2, 1], 3, 0.2, 0.2, 1.0, 4.0)
    res = pipeline_ref.denoise(w, cfg, z0, uncond, cond, steps, g=g, frozen_mask=frozen_mask,
                               frozen_latents=latents_all, frozen_steps=2)
    assert res["iters"] == [2, 1, 1, 0]
    assert (res["latents"] - lat_ref).abs().max() < 5e-3


def test_gligen_loop_with_ref_attention_matches_reference():
    """generate_gligen (LMD+ overall phase): fuser schedule, null-mask guidance pass, ref-attention loss, frozen blend"""
    cfg, w, r, md, z0, uncond, cond, table = _setup_pipeline(True)
    steps = 4
    g0 = torch.Generator().manual_seed(9)
    frozen_latents = torch.randn(steps + 1, 1, 4, 32, 32, generator=g0)
    frozen_latents[0] = z0
    frozen_mask = (torch.rand(32, 32, generator=g0) > 0.5).float()
    bboxes_flat = [(0.1, 0.2, 0.6, 0.7), (0.5, 0.4, 0.95, 0.9)]
    phrases = ["a cat", "a dog"]
    sg_bboxes = [[bboxes_flat[0]], [bboxes_flat[1]]]
    positions = [[2, 3], [6]]
    words = [3, 6]
    heads = 8
    refs = [[[{k: torch.softmax(3 * torch.randn(1, heads, 16 if k[0] == "mid" else 64, 1, generator=g0), dim=2)
               for k in KEYS} for _ in range(steps)]] for _ in range(2)]   # [obj][box][step][key]
    kw = dict(loss_scale=5, loss_threshold=0.01, max_iter=[2, 1], max_index_step=3, guidance_attn_keys=KEYS,
              fg_top_p=0.2, bg_top_p=0.2, fg_weight=1.0, bg_weight=4.0, use_ratio_based_loss=False,
              ref_ca_saved_attns=refs, ref_ca_word_token_only=True, word_token_indices=words, ref_ca_loss_weight=2.0,
              verbose=False)
    lat_ref, _ = r.pipelines.generate_gligen(
        md, frozen_latents, (uncond, cond), steps, bboxes_flat, phrases, gligen_scheduled_sampling_beta=0.5,
        frozen_steps=2, frozen_mask=frozen_mask, semantic_guidance=True, semantic_guidance_bboxes=sg_bboxes,
        semantic_guidance_object_positions=positions, semantic_guidance_kwargs=kw, show_progress=False)
    boxes = torch.zeros(1, 30, 4)
    boxes[0, :2] = torch.tensor(bboxes_flat)
    emb = torch.zeros(1, 30, 768)
    emb[0, :2] = table[:2]
    masks = torch.zeros(1, 30)
    masks[0, :2] = 1
    ref_maps = [[[{k: st[k][0, :, :, 0] for k in KEYS} for st in box] for box in obj] for obj in refs]
    g = pipeline_ref.GuidanceCfg(sg_bboxes, positions, KEYS, 5, 0.01, [2, 1], 3, 0.2, 0.2, 1.0, 4.0, ref_maps, words,
                                 2.0, True)
    res = pipeline_ref.denoise(w, cfg, z0, uncond, cond, steps, g=g, frozen_mask=frozen_mask,
                               frozen_latents=frozen_latents, frozen_steps=2,
                               gligen=dict(boxes=boxes, masks=masks, positive_embeddings=emb), gligen_beta=0.5)
    assert res["iters"] == [2, 1, 1, 0]
    assert (res["latents"] - lat_ref).abs().max() < 5e-3


@pytest.mark.parametrize("seed,smooth", [(0, True), (1, True), (2, False)])
def test_boxdiff_loss_and_grad_match_reference(seed, smooth):
    """utils/boxdiff.py compute_ca_loss_boxdiff (no reference-attention term) vs oracle/boxdiff_ref.py: loss and the
    gradient with respect to every saved map (groundwork for SURVEY section 8 row a14)"""
    from oracle import boxdiff_ref
    r = ref_loader.load()
    g = torch.Generator().manual_seed(100 + seed)
    heads, n, T = 8, 256, 77
    keys = [("down", 2, 0, 0), ("down", 2, 1, 0), ("up", 1, 0, 0), ("up", 1, 1, 0), ("up", 1, 2, 0)]
    bboxes = [[(0.1, 0.2, 0.6, 0.7)], [(0.5, 0.4, 0.95, 0.9), (0.0, 0.0, 0.3, 0.3)]]
    positions = [[2, 3], [6]]
    maps = {k: torch.softmax(2 * torch.randn(heads, n, T, generator=g), dim=-1) for k in keys}

    ours_in = {k: v.clone().requires_grad_(True) for k, v in maps.items()}
    ours = boxdiff_ref.boxdiff_loss(ours_in, bboxes, positions, keys, smooth_attentions=smooth)
    g_ours = torch.autograd.grad(ours, [ours_in[k] for k in keys])

    ref_in = {k: v.clone()[None].requires_grad_(True) for k, v in maps.items()}       # [1, heads, n, T]
    ref = r.boxdiff.compute_ca_loss_boxdiff(ref_in, bboxes, positions, keys, smooth_attentions=smooth)
    g_ref = torch.autograd.grad(ref, [ref_in[k] for k in keys])
    assert abs(float(ours) - float(ref)) < 1e-5 * max(1.0, abs(float(ref))), (float(ours), float(ref))
    for a, b in zip(g_ours, g_ref):
        assert (a - b[0]).abs().max() < 1e-6 + 1e-4 * b.abs().max()
    # update rule (boxdiff.py:228-232)
    z, gr = torch.randn(1, 4, 8, 8, generator=g), torch.randn(1, 4, 8, 8, generator=g)
    for index in (0, 7, 24):
        scale = (1.0 + (0.5 - 1.0) * index / 49) ** 0.5
        assert torch.allclose(boxdiff_ref.boxdiff_update(z, gr, index, 50), z - 20 * scale / 10 * gr)


def test_gligen_loop_fast_schedule_matches_reference():
    """generate_gligen with the thinned timestep list and per-step DDIM step size (fast_after_steps, fast_rate,
    dynamic_num_inference_steps; models/pipelines.py:358-362,439-440)"""
    cfg, w, r, md, z0, uncond, cond, table = _setup_pipeline(True)
    steps = 8
    bboxes_flat = [(0.1, 0.2, 0.6, 0.7)]
    lat_ref, _ = r.pipelines.generate_gligen(
        md, z0, (uncond, cond), steps, bboxes_flat, ["a cat"], gligen_scheduled_sampling_beta=0.5,
        semantic_guidance=False, show_progress=False, fast_after_steps=3, fast_rate=2, dynamic_num_inference_steps=True)
    boxes = torch.zeros(1, 30, 4)
    boxes[0, :1] = torch.tensor(bboxes_flat)
    emb = torch.zeros(1, 30, 768)
    emb[0, :1] = table[:1]
    masks = torch.zeros(1, 30)
    masks[0, :1] = 1
    res = pipeline_ref.denoise(w, cfg, z0, uncond, cond, steps, gligen=dict(boxes=boxes, masks=masks,
                                                                           positive_embeddings=emb),
                               gligen_beta=0.5, fast_after_steps=3, fast_rate=2, dynamic_num_inference_steps=True)
    assert res["latents_all"].shape[0] == 4            # initial + the three steps before the fast part
    assert (res["latents"] - lat_ref).abs().max() < 5e-3


def test_fast_schedule_matches_reference():
    """utils/schedule.py (get_fast_schedule, dynamically_adjust_inference_steps) vs pipelines.DDIMSchedule"""
    import lgd_b200  # noqa: F401
    from lgd_b200.pipelines import DDIMSchedule
    r = ref_loader.load()
    import importlib
    ref_sched = importlib.import_module("utils.schedule")

    class _Cfg:
        num_train_timesteps = 1000

    class _S:
        config = _Cfg()

    for steps in (10, 20, 50):
        for fast_after in (0, 3, steps // 2, steps - 2, steps - 1, steps + 5):
            for rate in (2, 3):
                ours = DDIMSchedule()
                ours.set_timesteps(steps)
                ref_ts = ref_sched.get_fast_schedule(torch.from_numpy(ours.timesteps.copy()), fast_after, rate)
                ours.apply_fast_schedule(fast_after, rate)
                assert ours.timesteps.tolist() == ref_ts.tolist()
                s = _S()
                s.timesteps = ref_ts
                for index, t in enumerate(ref_ts.tolist()):
                    import warnings
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        ref_sched.dynamically_adjust_inference_steps(s, index, t)
                    ours.adjust(index, t)
                    assert ours.num_inference_steps == s.num_inference_steps
    assert r is not None


def test_callshape_loop_matches_reference_pipeline():
    """oracle/callshape_ref.generate (the loop used to test the B200UNetAdapter on the GPU box) drives the UNMODIFIED
    reference UNet module through the reference's own call shape and must reproduce the unmodified
    pipelines.generate_gligen: same iteration counts, same latents, same saved maps"""
    from oracle import callshape_ref
    cfg, w, r, md, z0, uncond, cond, table = _setup_pipeline(True)
    steps = 4
    bboxes_flat = [(0.1, 0.2, 0.6, 0.7), (0.5, 0.4, 0.95, 0.9)]
    sg_bboxes = [[bboxes_flat[0]], [bboxes_flat[1]]]
    positions = [[2, 3], [6]]
    kw = dict(loss_scale=5, loss_threshold=0.01, max_iter=[2, 1], max_index_step=3, guidance_attn_keys=KEYS,
              fg_top_p=0.2, bg_top_p=0.2, fg_weight=1.0, bg_weight=4.0, use_ratio_based_loss=False, verbose=False)
    lat_ref, _, saved_ref = r.pipelines.generate_gligen(
        md, z0, (uncond, cond), steps, bboxes_flat, ["a cat", "a dog"], gligen_scheduled_sampling_beta=0.5,
        semantic_guidance=True, semantic_guidance_bboxes=sg_bboxes, semantic_guidance_object_positions=positions,
        semantic_guidance_kwargs=kw, show_progress=False, return_saved_cross_attn=True,
        saved_cross_attn_keys=[("down", 2, 1, 0)] + KEYS, return_cond_ca_only=True, return_token_ca_only=3)
    boxes = torch.zeros(1, 30, 4)
    boxes[0, :2] = torch.tensor(bboxes_flat)
    emb = torch.zeros(1, 30, 768)
    emb[0, :2] = table[:2]
    masks = torch.zeros(1, 30)
    masks[0, :2] = 1
    g = pipeline_ref.GuidanceCfg(sg_bboxes, positions, KEYS, 5, 0.01, [2, 1], 3, 0.2, 0.2, 1.0, 4.0)
    res = callshape_ref.generate(md.unet, z0, uncond, cond, steps, g=g,
                                 gligen=dict(boxes=boxes, masks=masks, positive_embeddings=emb), gligen_beta=0.5,
                                 saved_cross_attn_keys=[("down", 2, 1, 0)] + KEYS, return_token_ca_only=3,
                                 fuser_types=(r.attention.GatedSelfAttentionDense,))
    assert res["iters"] == [2, 1, 1, 0]
    assert (res["latents"] - lat_ref).abs().max() < 5e-3
    for s_ref, s in zip(saved_ref, res["saved"]):
        assert set(s_ref) == set(s)
        for k in s_ref:
            assert s_ref[k].shape == s[k].shape and (s_ref[k] - s[k]).abs().max() < 1e-3
    # and the CPU oracle behind the same call shape (what the GPU-box adapter test compares against)
    res2 = callshape_ref.generate(callshape_ref.OracleUNet(w, cfg), z0, uncond, cond, steps, g=g,
                                  gligen=dict(boxes=boxes, masks=masks, positive_embeddings=emb), gligen_beta=0.5)
    assert res2["iters"] == [2, 1, 1, 0] and (res2["latents"] - lat_ref).abs().max() < 5e-3


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_ratio_based_loss_matches_reference(seed):
    """compute_ca_lossv3 with its DEFAULT use_ratio_based_loss=True (what generation/backward_guidance.py runs,
    utils/guidance.py:122-128): loss and gradient w.r.t. every saved map vs oracle ca_loss(use_ratio_based_loss=True)"""
    import warnings
    r = ref_loader.load()
    saved, bboxes, positions, words, _ = _random_case(seed, with_ref=False)
    ref_in = {k: v.clone().requires_grad_(True) for k, v in saved.items()}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = r.guidance.compute_ca_lossv3(ref_in, bboxes, positions, KEYS, ref_ca_saved_attns=None,
                                           word_token_indices=words, ref_ca_loss_weight=0.5, verbose=False)
    g_ref = torch.autograd.grad(ref, [ref_in[k] for k in KEYS])
    ours_in = {k: v[0].clone().requires_grad_(True) for k, v in saved.items()}
    ours = guidance_ref.ca_loss(ours_in, bboxes, positions, KEYS, use_ratio_based_loss=True)
    g_ours = torch.autograd.grad(ours, [ours_in[k] for k in KEYS])
    assert abs(float(ours) - float(ref)) < 1e-6 * max(1.0, abs(float(ref)))
    for a, b in zip(g_ours, g_ref):
        assert (a - b[0]).abs().max() < 1e-7 + 1e-4 * b.abs().max()


def test_boxdiff_loop_matches_reference():
    """generate_semantic_guidance(use_boxdiff=True) (generation/boxdiff.py:126-137 -> utils/boxdiff.py:190-259): one
    BoxDiff step per denoising step with the sqrt step schedule, vs oracle pipeline_ref.denoise(boxdiff=...)"""
    cfg, w, r, md, z0, uncond, cond, _ = _setup_pipeline(False)
    keys = [("down", 2, 0, 0), ("down", 2, 1, 0), ("up", 1, 0, 0), ("up", 1, 1, 0), ("up", 1, 2, 0)]
    bboxes = [[(0.1, 0.2, 0.6, 0.7)], [(0.5, 0.4, 0.95, 0.9), (0.0, 0.0, 0.4, 0.4)]]
    positions = [[2, 3], [6]]
    steps = 4
    kw = dict(max_index_step=3, ref_ca_word_token_only=True, ref_ca_last_token_only=True, ref_ca_saved_attns=None,
              word_token_indices=[3, 6], guidance_attn_keys=keys, ref_ca_loss_weight=0.0, verbose=False)
    lat_ref, _ = r.pipelines.generate_semantic_guidance(
        md, z0, (torch.cat([uncond, cond]), uncond, cond), steps, bboxes, ["a cat", "a dog"], positions,
        semantic_guidance_kwargs=kw, use_boxdiff=True, show_progress=False)
    res = pipeline_ref.denoise(w, cfg, z0, uncond, cond, steps,
                               boxdiff=dict(bboxes=bboxes, object_positions=positions, keys=keys, max_index_step=3))
    assert len(res["boxdiff_losses"]) == 3
    assert (res["latents"] - lat_ref).abs().max() < 5e-3


# ---------------------------------------------------------------------------------------------------------------------
# mask refinement around the SAM network: lgd_b200.mask_refine vs the UNMODIFIED models/sam.py, both
# driven by the same synthetic "SAM network" (three candidate masks + predicted IoUs as a function of the prompt)
def _fake_prompt(input_boxes, input_points):
    if input_boxes is not None:          # [[[4]]] from sam_refine_boxes, [[4]] from sam_refine_attn (models/sam.py:141-142)
        b = input_boxes[0]
        return "box", (b if np.ndim(b) == 1 else b[0])
    return "point", input_points[0][0]


def _fake_sam_candidates(prompt_kind, prompt, size=512, seed=0):
    """deterministic stand-in for facebook/sam-vit-base: candidates grown from the prompt (whole / part / over-grown
    object, the usual three granularities), scores from a seeded generator"""
    rng = np.random.RandomState(seed + int(sum(np.ravel(prompt))) % 9973)
    yy, xx = np.mgrid[0:size, 0:size]
    if prompt_kind == "box":
        x0, y0, x1, y1 = [float(v) for v in prompt]
    else:
        px, py = [float(v) for v in prompt]
        w, h = rng.uniform(60, 200), rng.uniform(60, 200)
        x0, y0, x1, y1 = px - w / 2, py - h / 2, px + w / 2, py + h / 2
    cx, cy, rx, ry = (x0 + x1) / 2, (y0 + y1) / 2, max((x1 - x0) / 2, 2.0), max((y1 - y0) / 2, 2.0)
    ell = lambda s: (((xx - cx) / (rx * s)) ** 2 + ((yy - cy) / (ry * s)) ** 2) <= 1.0
    masks = np.stack([ell(0.9), ell(0.45) & (yy < cy), ell(rng.uniform(1.3, 2.5))])
    scores = rng.uniform(0.7, 1.0, size=3).astype(np.float32)
    return masks, scores


class _FakeSamInputs(dict):
    def to(self, device):
        return self


class _FakeSamProcessor:
    def __init__(self, seed):
        self.seed = seed
        self.image_processor = self

    def __call__(self, image, input_points=None, input_boxes=None, return_tensors="pt"):
        kind, prompt = _fake_prompt(input_boxes, input_points)
        if isinstance(image, list):                      # sam_refine_boxes hands over a list of images
            image = image[0]
        masks, scores = _fake_sam_candidates(kind, prompt, size=np.asarray(image).shape[0], seed=self.seed)
        return _FakeSamInputs(masks=torch.from_numpy(masks), scores=torch.from_numpy(scores),
                              original_sizes=torch.tensor([[masks.shape[1], masks.shape[2]]]),
                              reshaped_input_sizes=torch.tensor([[1024, 1024]]))

    def post_process_masks(self, pred_masks, original_sizes, reshaped_input_sizes):
        return [pred_masks[0] > 0]                       # [n_prompts = 1, 3, h, w] bool per image


class _FakeSamModel:
    def __call__(self, masks, scores, original_sizes, reshaped_input_sizes):
        import types
        return types.SimpleNamespace(pred_masks=(masks.float() * 2 - 1)[None, None], iou_scores=scores[None, None])


def _fake_predict(seed):
    def predict(image, input_boxes=None, input_points=None):
        kind, prompt = _fake_prompt(input_boxes, input_points)
        return _fake_sam_candidates(kind, prompt, size=np.asarray(image).shape[0], seed=seed)
    return predict


def _ref_sam():
    ref_loader.load()
    from models import sam as rsam
    rsam.torch_device = "cpu"
    return rsam


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_mask_refine_box_matches_reference_sam(seed):
    from lgd_b200 import mask_refine as MR
    rsam = _ref_sam()
    rng = np.random.RandomState(seed)
    image = rng.randint(0, 255, size=(512, 512, 3)).astype(np.uint8)
    md = dict(sam_model=_FakeSamModel(), sam_processor=_FakeSamProcessor(seed))
    for _ in range(6):
        x0, y0 = rng.uniform(0, 0.6, size=2)
        box = (x0, y0, x0 + rng.uniform(0.08, 0.4), y0 + rng.uniform(0.08, 0.4))
        conf_th, iou_th = rng.choice([0.85, 0.8, 0.95]), rng.choice([0.2, 0.25, 0.6])
        m_ref, c_ref = rsam.sam_refine_box(image, box, model_dict=md, height=512, width=512, H=64, W=64,
                                           discourage_mask_below_confidence=conf_th,
                                           discourage_mask_below_coarse_iou=iou_th, verbose=False)
        m, c = MR.refine_box(_fake_predict(seed), image, box, 512, 512, 64, 64, discourage_mask_below_confidence=conf_th,
                             discourage_mask_below_coarse_iou=iou_th)
        assert m.dtype == np.bool_ and m.shape == (64, 64)
        assert np.array_equal(m, m_ref) and float(c) == float(c_ref)


@pytest.mark.parametrize("use_box_input", [False, True])
@pytest.mark.parametrize("side", [16, 64])
def test_mask_refine_attn_matches_reference_sam(use_box_input, side):
    from lgd_b200 import mask_refine as MR
    rsam = _ref_sam()
    rng = np.random.RandomState(10 + side + int(use_box_input))
    image = rng.randint(0, 255, size=(512, 512, 3)).astype(np.uint8)
    md = dict(sam_model=_FakeSamModel(), sam_processor=_FakeSamProcessor(5))
    sigma = MR.GAUSSIAN_SIGMA_BOX_INPUT if use_box_input else MR.GAUSSIAN_SIGMA_POINT_INPUT
    for _ in range(6):
        # a token-attention map: a blob plus noise, like utils/attn.py get_token_attnv2 hands over
        yy, xx = np.mgrid[0:side, 0:side] / side
        cx, cy, s = rng.uniform(0.25, 0.75), rng.uniform(0.25, 0.75), rng.uniform(0.08, 0.2)
        attn = (np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * s * s)) + 0.05 * rng.rand(side, side)).astype(np.float32)
        kw = dict(use_box_input=use_box_input, gaussian_sigma=sigma, mask_th_for_box=0.05, n_erode_dilate_mask_for_box=1,
                  mask_th_for_point=0.25, discourage_mask_below_confidence=0.85, discourage_mask_below_coarse_iou=0.25)
        m_ref, c_ref = rsam.sam_refine_attn(image, attn.copy(), model_dict=md, height=512, width=512, H=64, W=64,
                                            verbose=False, **kw)
        m, c = MR.refine_attn(_fake_predict(5), image, attn.copy(), 512, 512, 64, 64, **kw)
        assert np.array_equal(m, m_ref) and float(c) == float(c_ref)
        mb, prompt = MR.attn_prompt(attn, 512, 512, use_box_input, sigma)
        assert mb.shape == (side, side) and (("input_boxes" in prompt) == use_box_input)


def test_select_mask_rule_matches_reference():
    from lgd_b200 import mask_refine as MR
    rsam = _ref_sam()
    rng = np.random.RandomState(0)
    for _ in range(50):
        masks = rng.rand(3, 32, 32) > rng.uniform(0.2, 0.9, size=(3, 1, 1))
        conf = rng.uniform(0.6, 1.0, size=3)
        ious = rng.uniform(0.0, 0.6, size=3) if rng.rand() < 0.7 else None
        a, ca = rsam.select_mask(masks, conf, coarse_ious=ious, discourage_mask_below_confidence=0.85,
                                 discourage_mask_below_coarse_iou=0.2)
        b, cb = MR.select_mask(masks, conf, ious, 0.85, 0.2)
        assert np.array_equal(a, b) and ca == cb
