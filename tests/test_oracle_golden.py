"""Oracle vs the committed golden vectors (tests/golden/*.npz, produced BY the unmodified reference with
oracle/make_goldens.py).  Runs anywhere - including the GPU box where /root/reference does not exist."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import guidance_ref, pipeline_ref, unet_ref  # noqa: E402

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KEYS = [("mid", 0, 0, 0), ("up", 1, 0, 0), ("up", 1, 1, 0), ("up", 1, 2, 0)]


def test_loss_and_gradient_goldens():
    import test_oracle_vs_reference as T
    gold = np.load(os.path.join(G, "loss_reference.npz"))
    for seed in (0, 1, 2, 3):
        for with_ref in (False, True):
            saved, bboxes, positions, words, refs = T._random_case(seed, with_ref=with_ref)
            one = {k: v[0].numpy() for k, v in saved.items()}
            ref_maps = None
            if refs is not None:
                ref_maps = [[{k: box[0][k][0, :, :, 0].numpy() for k in KEYS} for box in obj] for obj in refs]
            L, grads = guidance_ref.ca_loss_and_grad(one, bboxes, positions, KEYS, 0.2, 0.2, 1.0, 4.0, ref_maps, words,
                                                     2.0, True)
            tag = f"s{seed}_r{int(with_ref)}"
            assert abs(L - float(gold[tag + "_loss"])) < 5e-6 * max(1.0, abs(L))
            for k in KEYS:
                ks = "_".join(map(str, k))
                np.testing.assert_allclose(grads[k].sum(axis=1), gold[tag + "_g_" + ks], rtol=3e-4, atol=3e-7)
                assert abs(np.abs(grads[k]).sum() - float(gold[tag + "_gabs_" + ks])) < 3e-4 * float(gold[tag + "_gabs_" + ks])


def test_unet_goldens():
    gold = np.load(os.path.join(G, "unet_reference.npz"))
    for gl in (False, True):
        cfg = unet_ref.UNetConfig.tiny(gligen=gl)
        w = unet_ref.make_weights(cfg, seed=0)
        g = torch.Generator().manual_seed(1)
        x = torch.randn(2, 4, 16, 16, generator=g)
        ctx = torch.randn(2, 77, 768, generator=g)
        glin = None
        if gl:
            glin = dict(boxes=torch.rand(2, 30, 4, generator=g), masks=(torch.rand(2, 30, generator=g) > 0.8).float(),
                        positive_embeddings=torch.randn(2, 30, 768, generator=g))
        with torch.no_grad():
            saved = {}
            eps = unet_ref.unet_forward(w, cfg, x, 481, ctx, gligen=glin, saved=saved)
        assert np.abs(eps.numpy() - gold[f"eps_gligen{int(gl)}"]).max() < 3e-5
        assert np.abs(saved[("mid", 0, 0, 0)].numpy() - gold[f"mid_probs_gligen{int(gl)}"]).max() < 1e-4


def test_pipeline_golden():
    gold = np.load(os.path.join(G, "pipeline_reference.npz"))
    cfg = unet_ref.UNetConfig.tiny()
    w = unet_ref.make_weights(cfg, seed=0)
    g = torch.Generator().manual_seed(5)
    z0 = torch.randn(1, 4, 32, 32, generator=g)
    uncond = torch.randn(1, 77, 768, generator=g)
    cond = torch.randn(1, 77, 768, generator=g)
    gc = pipeline_ref.GuidanceCfg([[(0.1, 0.2, 0.6, 0.7)], [(0.5, 0.4, 0.95, 0.9)]], [[2, 3], [6]], KEYS, 30, 0.2,
                                  [2, 1, 1], 3, 0.2, 0.2, 1.0, 4.0)
    res = pipeline_ref.denoise(w, cfg, z0, uncond, cond, 4, g=gc)
    assert res["iters"] == [2, 1, 1, 0]
    assert np.abs(res["latents_all"].numpy() - gold["semantic_latents_all"]).max() < 5e-3


def test_boxdiff_goldens():
    """oracle/boxdiff_ref.py vs the reference's compute_ca_loss_boxdiff outputs frozen in boxdiff_reference.npz"""
    from oracle import boxdiff_ref
    from oracle.boxdiff_ref import BOXDIFF_KEYS, boxdiff_inputs
    gold = np.load(os.path.join(G, "boxdiff_reference.npz"))
    for seed in (0, 1, 2):
        maps, bboxes, positions = boxdiff_inputs(seed)
        leaf = {k: v.clone().requires_grad_(True) for k, v in maps.items()}
        L = boxdiff_ref.boxdiff_loss(leaf, bboxes, positions, BOXDIFF_KEYS)
        grads = torch.autograd.grad(L, [leaf[k] for k in BOXDIFF_KEYS])
        assert abs(float(L) - float(gold[f"s{seed}_loss"])) < 1e-5 * max(1.0, abs(float(L)))
        for k, gr in zip(BOXDIFF_KEYS, grads):
            ks = "_".join(map(str, k))
            np.testing.assert_allclose(gr.sum(dim=1).numpy(), gold[f"s{seed}_g_{ks}"], rtol=1e-3, atol=1e-7)
            assert abs(float(gr.abs().sum()) - float(gold[f"s{seed}_gabs_{ks}"])) < 1e-3 * float(gold[f"s{seed}_gabs_{ks}"])
