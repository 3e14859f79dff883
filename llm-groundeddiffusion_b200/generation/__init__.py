"""generation.* plug-ins with the reference's surface: module-level `version` and
`run(spec, bg_seed, fg_seed_start, **kwargs) -> EasyDict(image=..., so_img_list=...)` (generate.py:130-154,323-386)."""
