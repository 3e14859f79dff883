"""LMD (attention guidance per box and overall, no GLIGEN) - generation/lmd.py of the reference; keyword surface and
defaults of generation/lmd.py:215-256.  Deviation (stated in DESIGN.md): the SAM refinement step that the reference runs
between the phases is supplied by the environment (`env.refine_mask`); without SAM weights it is the box raster."""
from . import common
from .common import DEFAULT_OVERALL_NEGATIVE_PROMPT, DEFAULT_SO_NEGATIVE_PROMPT

version = "lmd"
_MAX_ITER = [4] * 5 + [3] * 5 + [2] * 5 + [2] * 5 + [1] * 10
# generation/lmd.py:36-37: the word-token attention handed to SAM is averaged over the steps from this index on
attn_aggregation_step_start = 10


def run_batch(specs, bg_seeds, fg_seed_starts, overall_prompt_overrides=None, frozen_step_ratio=0.5,
              num_inference_steps=50, loss_scale=5, loss_threshold=5.0, max_iter=_MAX_ITER, max_index_step=30,
              overall_loss_scale=5, overall_loss_threshold=5.0, overall_max_iter=_MAX_ITER, overall_max_index_step=30,
              fg_top_p=0.2, bg_top_p=0.2, overall_fg_top_p=0.2, overall_bg_top_p=0.2, fg_weight=1.0, bg_weight=4.0,
              overall_fg_weight=1.0, overall_bg_weight=4.0, ref_ca_loss_weight=2.0, so_center_box=True,
              fg_blending_ratio=0.01, so_negative_prompt=DEFAULT_SO_NEGATIVE_PROMPT,
              overall_negative_prompt=DEFAULT_OVERALL_NEGATIVE_PROMPT, mask_th_for_point=0.25,
              so_horizontal_center_only=False, align_with_overall_bboxes=True, horizontal_shift_only=False,
              use_fast_schedule=False, so_vertical_placement="floor_padding", so_floor_padding=0.2, use_box_input=False,
              use_ref_ca=True, use_autocast=False, verbose=False, return_latents=False):
    so_g = dict(loss_scale=loss_scale, loss_threshold=loss_threshold, max_iter=max_iter, max_index_step=max_index_step,
                fg_top_p=fg_top_p, bg_top_p=bg_top_p, fg_weight=fg_weight, bg_weight=bg_weight)
    ov_g = dict(loss_scale=overall_loss_scale, loss_threshold=overall_loss_threshold, max_iter=overall_max_iter,
                max_index_step=overall_max_index_step, fg_top_p=overall_fg_top_p, bg_top_p=overall_bg_top_p,
                fg_weight=overall_fg_weight, bg_weight=overall_bg_weight)
    return common.layout_generation(
        specs, bg_seeds, fg_seed_starts, use_gligen=False, so_guidance=so_g, overall_guidance=ov_g,
        num_inference_steps=num_inference_steps, frozen_step_ratio=frozen_step_ratio, so_beta=0.0, overall_beta=0.0,
        so_center_box=so_center_box, so_horizontal_center_only=so_horizontal_center_only,
        so_vertical_placement=so_vertical_placement, so_floor_padding=so_floor_padding,
        fg_blending_ratio=fg_blending_ratio, align_with_overall_bboxes=align_with_overall_bboxes,
        horizontal_shift_only=horizontal_shift_only, use_ref_ca=use_ref_ca, ref_ca_loss_weight=ref_ca_loss_weight,
        so_negative_prompt=so_negative_prompt, overall_negative_prompt=overall_negative_prompt,
        overall_prompt_overrides=overall_prompt_overrides, return_latents=return_latents,
        use_fast_schedule=use_fast_schedule, sam_attn_start=attn_aggregation_step_start)


def run(spec, bg_seed=1, overall_prompt_override="", fg_seed_start=20, **kwargs):
    return run_batch([spec], [bg_seed], [fg_seed_start], overall_prompt_overrides=[overall_prompt_override],
                     **kwargs)[0]
