"""LMD+ (GLIGEN adapters + attention guidance + reference-attention transfer) - generation/lmd_plus.py of the reference.
Same keyword surface and defaults as generation/lmd_plus.py:193-228; `run_batch` is the B200 addition (B specs in
lock-step on one GPU)."""
from . import common
from .common import DEFAULT_OVERALL_NEGATIVE_PROMPT, DEFAULT_SO_NEGATIVE_PROMPT

version = "lmd_plus"
_MAX_ITER = [4] * 5 + [3] * 5 + [2] * 5 + [2] * 5 + [1] * 10


def run_batch(specs, bg_seeds, fg_seed_starts, overall_prompt_overrides=None, frozen_step_ratio=0.5,
              num_inference_steps=50, loss_scale=5, loss_threshold=5.0, max_iter=_MAX_ITER, max_index_step=0,
              overall_loss_scale=5, overall_loss_threshold=5.0, overall_max_iter=_MAX_ITER, overall_max_index_step=30,
              so_gligen_scheduled_sampling_beta=0.4, overall_gligen_scheduled_sampling_beta=0.4, overall_fg_top_p=0.2,
              overall_bg_top_p=0.2, overall_fg_weight=1.0, overall_bg_weight=4.0, ref_ca_loss_weight=2.0,
              so_center_box=False, fg_blending_ratio=0.1, so_negative_prompt=DEFAULT_SO_NEGATIVE_PROMPT,
              overall_negative_prompt=DEFAULT_OVERALL_NEGATIVE_PROMPT, so_horizontal_center_only=True,
              align_with_overall_bboxes=False, horizontal_shift_only=True, use_fast_schedule=False, use_ref_ca=True,
              use_autocast=True, verbose=False, return_latents=False):
    so_g = dict(loss_scale=loss_scale, loss_threshold=loss_threshold, max_iter=max_iter, max_index_step=max_index_step)
    ov_g = dict(loss_scale=overall_loss_scale, loss_threshold=overall_loss_threshold, max_iter=overall_max_iter,
                max_index_step=overall_max_index_step, fg_top_p=overall_fg_top_p, bg_top_p=overall_bg_top_p,
                fg_weight=overall_fg_weight, bg_weight=overall_bg_weight)
    return common.layout_generation(
        specs, bg_seeds, fg_seed_starts, use_gligen=True, so_guidance=so_g, overall_guidance=ov_g,
        num_inference_steps=num_inference_steps, frozen_step_ratio=frozen_step_ratio,
        so_beta=so_gligen_scheduled_sampling_beta, overall_beta=overall_gligen_scheduled_sampling_beta,
        so_center_box=so_center_box, so_horizontal_center_only=so_horizontal_center_only,
        so_vertical_placement="centered", so_floor_padding=None, fg_blending_ratio=fg_blending_ratio,
        align_with_overall_bboxes=align_with_overall_bboxes, horizontal_shift_only=horizontal_shift_only,
        use_ref_ca=use_ref_ca, ref_ca_loss_weight=ref_ca_loss_weight, so_negative_prompt=so_negative_prompt,
        overall_negative_prompt=overall_negative_prompt, overall_prompt_overrides=overall_prompt_overrides,
        return_latents=return_latents, use_fast_schedule=use_fast_schedule)


def run(spec, bg_seed=1, overall_prompt_override="", fg_seed_start=20, **kwargs):
    return run_batch([spec], [bg_seed], [fg_seed_start], overall_prompt_overrides=[overall_prompt_override],
                     **kwargs)[0]
