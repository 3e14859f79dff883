"""ctypes binding of libb200lmd.so (the C ABI declared in include/b200lmd.h).

There is no CPU fallback: if the shared library is missing, importing a symbol raises immediately.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200lmd.so")

_lib = None


class B200Error(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error(
                f"{LIB_PATH} not built - run `python -c 'import __graft_entry__ as g; g.build()'` (there is no "
                "CPU or PyTorch fallback for the B200 path)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.b200lmd_last_error.restype = ctypes.c_char_p
        # A/B switches between kernel generations for measurements / bisection (all default to the fastest correct
        # variant): B200_OPT_GEMM_V2=0, B200_OPT_GEMM_V3=0, B200_OPT_ATTN_V2=0, B200_OPT_FUSED_LOSS_STAGE=1
        for name in ("gemm_v2", "gemm_v3", "attn_v2", "fused_loss_stage"):
            v = os.environ.get("B200_OPT_" + name.upper())
            if v is not None:
                _lib.b200lmd_set_option(name.encode(), ctypes.c_int(int(v)))
    return _lib


_calls = 0


def check(rc):
    """every C-ABI op call funnels through here; each successful call launched >= 1 of our kernels"""
    global _calls
    if rc != 0:
        raise B200Error(lib().b200lmd_last_error().decode())
    _calls += 1


def launch_count():
    """lower bound on the number of our kernel launches so far (C-ABI op calls; several launch 2-4 kernels)"""
    return _calls


def ptr(t):
    """device pointer of a torch tensor (or None)"""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def cur_stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
