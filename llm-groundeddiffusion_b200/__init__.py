"""B200-native layout-grounded denoising path (LLM-grounded Diffusion hot path) - host side.

Python here only marshals tensors into the C ABI of libb200lmd.so (include/b200lmd.h); all arithmetic on the path runs
in hand-written sm_100a kernels.  Import as `lgd_b200`.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
