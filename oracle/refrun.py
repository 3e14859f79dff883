"""Drive the UNMODIFIED reference (via oracle/ref_loader.py) on CPU with seeded synthetic weights - build container only.

Used by tests/test_oracle_vs_reference.py and oracle/make_goldens.py to pin the oracle restatement against outputs of
the reference itself.  TEST INFRASTRUCTURE ONLY.
"""
import types

import torch

from . import ref_loader, unet_ref


class _FakeVAE:
    """the reference loops end in vae.decode (models/pipelines.py:233,461,591); images are not part of the parity"""

    def decode(self, z):
        return types.SimpleNamespace(sample=torch.zeros(z.shape[0], 3, 8, 8))


class _TokOut(dict):
    def to(self, device):
        return self


class FakeTokenizer:
    """tokenizer stand-in for prepare_gligen_condition (models/pipelines.py:303): carries phrase ids through"""

    def __init__(self, phrase_ids):
        self.phrase_ids = phrase_ids

    def __call__(self, phrases, padding=True, return_tensors="pt", **kw):
        return _TokOut(input_ids=torch.tensor([[self.phrase_ids[p]] for p in phrases]))


class FakeTextEncoder:
    """text_encoder(**inputs).pooler_output -> seeded phrase embedding table lookup"""

    def __init__(self, table):
        self.table = table

    def __call__(self, input_ids=None, **kw):
        return types.SimpleNamespace(pooler_output=self.table[input_ids[:, 0]])


def build_reference_unet(cfg: unet_ref.UNetConfig, w):
    r = ref_loader.load()
    m = r.unet_2d_condition.UNet2DConditionModel(**cfg.to_reference_kwargs())
    m.load_state_dict(w)
    m.eval()
    for p in m.parameters():
        p.requires_grad_(False)
    return m


def model_dict(cfg, w, tokenizer=None, text_encoder=None):
    r = ref_loader.load()
    from easydict import EasyDict
    from diffusers import DDIMScheduler
    return r, EasyDict(vae=_FakeVAE(), tokenizer=tokenizer, text_encoder=text_encoder,
                       unet=build_reference_unet(cfg, w), scheduler=DDIMScheduler(), dtype=torch.float32)
