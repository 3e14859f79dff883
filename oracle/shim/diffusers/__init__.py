"""Minimal stand-in for diffusers==0.18.0 (not installable offline) so that the UNMODIFIED reference files under
/root/reference/models can be imported for oracle validation in the build container.

TEST INFRASTRUCTURE ONLY.  Everything here restates the published diffusers-0.18.0 definitions that the reference
imports: ResnetBlock2D / Downsample2D / Upsample2D (models/resnet.py), Timesteps /
TimestepEmbedding (models/embeddings.py), DDIMScheduler (schedulers/scheduling_ddim.py) and the config/mixin plumbing.
Parity of THESE pieces against the real diffusers wheel is unpinned (no wheel offline).
"""
from .schedulers import DDIMScheduler  # noqa: F401


class _Unavailable:
    def __init__(self, *a, **k):
        raise RuntimeError("not available in the diffusers shim")


AutoencoderKL = DDIMInverseScheduler = DPMSolverMultistepScheduler = _Unavailable
