"""Import the UNMODIFIED reference (/root/reference) on CPU for oracle validation and golden-vector minting.

TEST INFRASTRUCTURE ONLY - exists only in the build container (the GPU box has no /root/reference); nothing on the
product path, in `-m gpu` tests, smoke() or bench.py may import this module.

What it does: puts /root/reference and oracle/shim (a stand-in for the missing diffusers 0.18 / easydict / matplotlib
packages) on sys.path, and neutralises the reference's hard-coded CUDA placement (utils/guidance.py:104,186,204,253,
utils/boxdiff.py:76,130, utils/utils.py:6) by giving those modules a `torch` proxy whose factory functions ignore
device="cuda" and by making Tensor.cuda() the identity.
"""
import os
import sys
import types

import torch

REFERENCE_ROOT = os.environ.get("LMD_REFERENCE_ROOT", "/root/reference")
SHIM_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


class _TorchProxy(types.ModuleType):
    """`torch` as seen by reference modules: identical, except device='cuda' requests land on the CPU."""

    def __init__(self):
        super().__init__("torch")

    def __getattr__(self, name):
        return getattr(torch, name)


def _cpu_factory(fn):
    def wrapped(*a, **k):
        if "device" in k and str(k["device"]).startswith("cuda"):
            k["device"] = "cpu"
        return fn(*a, **k)
    return wrapped


_loaded = {}


def load():
    """returns a namespace with the reference modules: models_pkg (unet etc.), guidance, boxdiff, attn, schedule, utils"""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    for p in (SHIM_ROOT, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    torch.Tensor.cuda = lambda self, *a, **k: self  # noqa: identity on CPU
    torch.nn.Module.cuda = lambda self, *a, **k: self  # utils/boxdiff.py:76 moves its smoothing module to the GPU

    proxy = _TorchProxy()
    for name in ("zeros", "ones", "tensor", "empty", "arange", "full", "zeros_like"):
        setattr(proxy, name, _cpu_factory(getattr(torch, name)))

    import utils as ref_utils  # reference package `utils`
    import utils.utils as ref_utils_utils
    ref_utils_utils.torch_device = "cpu"
    ref_utils_utils.torch = proxy
    # the reference does `from .utils import *`-style re-export through utils/__init__.py
    for k in ("torch_device",):
        if hasattr(ref_utils, k):
            setattr(ref_utils, k, "cpu")
    from utils import guidance, schedule, boxdiff, attn
    for m in (guidance, boxdiff, attn):
        m.torch = proxy

    from models import attention, attention_processor, transformer_2d, unet_2d_blocks, unet_2d_condition
    _loaded.update(dict(utils=ref_utils, guidance=guidance, schedule=schedule, boxdiff=boxdiff, attn=attn,
                        attention=attention, attention_processor=attention_processor, transformer_2d=transformer_2d,
                        unet_2d_blocks=unet_2d_blocks, unet_2d_condition=unet_2d_condition))
    try:
        from models import pipelines
        pipelines.torch_device = "cpu"
        _loaded["pipelines"] = pipelines
    except Exception as e:  # models/models.py needs transformers CLIP classes + diffusers AutoencoderKL names only
        _loaded["pipelines_error"] = repr(e)
    return types.SimpleNamespace(**_loaded)
