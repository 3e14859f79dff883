"""Mint the golden fixtures of the two-phase layout-grounded generation by running the UNMODIFIED reference plug-ins
(generation/lmd.py, generation/lmd_plus.py `run`) on CPU through oracle/refrun_lmd.py - build container only.
TEST INFRASTRUCTURE.  Usage (from the repo root):

    python -m oracle.make_goldens_layout config1        # BASELINE config 1: LMD, SD1.5 widths, fp32 CPU, 10 steps,
                                                        # 2 boxes, bg_seed 0, fg_seed_start 20  (~1 h on 8 cores)
    python -m oracle.make_goldens_layout lmdplus_tiny   # LMD+ (GLIGEN) at the small topology, 2 specs, 6 steps
    python -m oracle.make_goldens_layout lmdplus_sd15   # LMD+ at SD1.5+GLIGEN widths, 2 boxes, 10 steps (~45 min)
    python -m oracle.make_goldens_layout lmd_tiny       # LMD at the small topology (fast-schedule variant)

Stated deviations from a stock run: synthetic seeded weights / text embeddings (no checkpoints
offline), SAM replaced by the box raster, and `generation.lmd.attn_aggregation_step_start` lowered from 10 to 5 for the
10-step run (with the stock value the reference stacks an empty list at utils/attn.py:18 and raises).
Fixtures: tests/golden/layout_<name>.npz.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import refrun_lmd, unet_ref  # noqa: E402

DEER_BEAR = dict(prompt="A realistic image of a white deer and a gray bear in an empty factory scene",
                 gen_boxes=[("a white deer", [74, 177, 183, 235]), ("a gray bear", [314, 193, 189, 216])],
                 bg_prompt="A realistic image of an empty factory scene", extra_neg_prompt="")
CUBES = dict(prompt="In an indoor scene, a blue cube directly above a red cube with a vase on the left of them",
             gen_boxes=[("a blue cube", [202, 120, 110, 110]), ("a red cube", [202, 230, 110, 110]),
                        ("a vase", [62, 190, 80, 150])], bg_prompt="An indoor scene", extra_neg_prompt="")
TWO_CATS = dict(prompt="two cats and a dog on the grass",
                gen_boxes=[("a cat", [40, 250, 150, 180]), ("a cat", [300, 260, 160, 170]), ("a dog", [180, 120, 140, 150])],
                bg_prompt="A photo of the grass", extra_neg_prompt="people")

CASES = {
    # name: (method, unet config, weight seed, [(spec, bg_seed, fg_seed_start)], run kwargs, attn_aggregation_step_start)
    "config1": ("lmd", "sd15", 0, [(DEER_BEAR, 0, 20)], dict(num_inference_steps=10), 5),
    "lmd_tiny": ("lmd", "tiny", 0, [(CUBES, 3, 20)],
                 dict(num_inference_steps=8, use_fast_schedule=True, overall_max_index_step=5, max_index_step=5,
                      max_iter=[2, 1], overall_max_iter=[2, 1]), 4),
    "lmdplus_tiny": ("lmd_plus", "tiny_gligen", 0, [(DEER_BEAR, 0, 20), (TWO_CATS, 7, 30)],
                     dict(num_inference_steps=6, overall_max_iter=[2, 2, 1], overall_max_index_step=4), None),
    # the benchmarked function at the benchmark's widths: LMD+ (SD1.4/1.5 + GLIGEN shapes), 2 boxes, 10 steps
    "lmdplus_sd15": ("lmd_plus", "sd15_gligen", 0, [(DEER_BEAR, 0, 20)], dict(num_inference_steps=10), None),
}


def config_of(name):
    return {"sd15": unet_ref.UNetConfig.sd15(), "sd15_gligen": unet_ref.UNetConfig.sd15(gligen=True),
            "tiny": unet_ref.UNetConfig.tiny(), "tiny_gligen": unet_ref.UNetConfig.tiny(gligen=True)}[name]


def mint(name):
    method, cfg_name, wseed, runs, kw, agg = CASES[name]
    cfg = config_of(cfg_name)
    w = unet_ref.make_weights(cfg, seed=wseed)
    out = {"meta": json.dumps(dict(method=method, cfg=cfg_name, weight_seed=wseed, run_kwargs=kw,
                                   attn_aggregation_step_start=agg, n_runs=len(runs),
                                   specs=[r[0] for r in runs], seeds=[[r[1], r[2]] for r in runs]))}
    for i, (spec, bg, fg) in enumerate(runs):
        t0 = time.time()
        _, rec = refrun_lmd.run_reference(method, cfg, w, spec, bg, fg, kw, attn_aggregation_step_start=agg)
        print(f"[{name}] run {i}: {time.time() - t0:.0f} s, {len(rec.generations)} generations", flush=True)
        for j, g in enumerate(rec.generations):
            p = f"r{i}_g{j}_"
            out[p + "kind"] = np.array(g["kind"])
            out[p + "latents"] = g["latents"].float().numpy()
            out[p + "losses"] = np.array(g["losses"], dtype=np.float64)
            out[p + "iters"] = np.array(g.get("iters", []), dtype=np.int64)
        out[f"r{i}_n_gen"] = np.array(len(rec.generations))
        out[f"r{i}_masks"] = np.stack(rec.masks) if rec.masks else np.zeros((0, 1, 1), bool)
        out[f"r{i}_sam_inputs"] = np.stack(rec.sam_inputs) if rec.sam_inputs else np.zeros((0, 1, 1), np.float32)
    path = os.path.join(ROOT, "tests", "golden", f"layout_{name}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    for n in sys.argv[1:]:
        mint(n)
