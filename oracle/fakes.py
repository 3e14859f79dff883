"""TEST INFRASTRUCTURE (never on the product path): offline stand-ins for the third-party models around the path -
CLIP tokenizer / text encoder, VAE - with the interfaces the reference touches (models/models.py:63-89,
utils/guidance.py:10-30, models/pipelines.py:117-127,303-304).  No CLIP vocabulary or SD weights exist offline.

The SAME objects drive (a) the unmodified reference on CPU (oracle/refrun_lmd.py, build container) and (b) the B200
path on the GPU box through lgd_b200.env.ReferenceEnv, so both sides see identical text embeddings and token indices.
"""
import hashlib
import types

import numpy as np
import torch


def _seed_of(s):
    return int.from_bytes(hashlib.sha256(s.encode()).digest()[:4], "little")


class _Batch(dict):
    """BatchEncoding stand-in: attribute + key access, .to(device)"""
    __getattr__ = dict.get

    def to(self, device):
        return _Batch({k: (v.to(device) if torch.is_tensor(v) else v) for k, v in self.items()})


class WordTokenizer:
    """<bos> word word ... <eos>, one token per whitespace-separated word (commas split off), growing vocabulary"""
    model_max_length = 77
    bos_token, eos_token = "<bos>", "<eos>"

    def __init__(self):
        self.inv = [self.bos_token, self.eos_token]
        self.vocab = {t: i for i, t in enumerate(self.inv)}
        self.eos_token_id = 1

    def _id(self, tok):
        if tok not in self.vocab:
            self.vocab[tok] = len(self.inv)
            self.inv.append(tok)
        return self.vocab[tok]

    def words(self, text):
        return text.replace(",", " ,").split()

    def __call__(self, prompts, padding="do_not_pad", max_length=None, truncation=False, return_tensors="pt", **kw):
        if isinstance(prompts, str):
            prompts = [prompts]
        rows = [[0] + [self._id(w) for w in self.words(p)] + [1] for p in prompts]
        if padding == "max_length":
            L = max_length or self.model_max_length
            rows = [(r[:L - 1] + [1] if len(r) > L else r) + [1] * (L - len(r)) for r in rows]
        elif padding is True or padding == "longest":
            L = max(len(r) for r in rows)
            rows = [r + [1] * (L - len(r)) for r in rows]
        if return_tensors == "np":
            ids = np.array(rows, dtype=np.int64) if len({len(r) for r in rows}) == 1 else \
                np.array([np.array(r) for r in rows], dtype=object)
        else:
            ids = torch.tensor(rows, dtype=torch.long)
        return _Batch(input_ids=ids)

    def _convert_id_to_token(self, i):
        return self.inv[int(i)]

    def text_of(self, ids):
        """canonical string of an id row (bos / eos / padding dropped)"""
        return " ".join(self.inv[int(i)] for i in ids if int(i) > 1)


class FakeTextEncoder(torch.nn.Module):
    """text_encoder(input_ids)[0] -> [B, L, ctx] seeded by the canonical string of the row;
    text_encoder(**inputs).pooler_output -> [B, ctx] (GLIGEN phrase embeddings, models/pipelines.py:303-304)"""

    def __init__(self, tokenizer, ctx_dim=768):
        super().__init__()
        self.tok, self.ctx_dim = tokenizer, ctx_dim
        self._dummy = torch.nn.Parameter(torch.zeros(1), requires_grad=False)
        self._cache = {}

    def embed(self, text, rows):
        key = (text, rows)
        if key not in self._cache:
            g = torch.Generator().manual_seed(_seed_of(text))
            self._cache[key] = torch.randn(rows, self.ctx_dim, generator=g)
        return self._cache[key]

    def forward(self, input_ids=None, **kw):
        ids = input_ids.cpu()
        hidden = torch.stack([self.embed("seq:" + self.tok.text_of(r), ids.shape[1]) for r in ids.tolist()])
        pooled = torch.stack([self.embed("pool:" + self.tok.text_of(r), 1)[0] for r in ids.tolist()])
        dev = self._dummy.device
        out = types.SimpleNamespace(last_hidden_state=hidden.to(dev), pooler_output=pooled.to(dev))
        return _Out(out)


class _Out:
    """indexable like a transformers ModelOutput: out[0] = last_hidden_state"""

    def __init__(self, ns):
        self.last_hidden_state, self.pooler_output = ns.last_hidden_state, ns.pooler_output

    def __getitem__(self, i):
        return (self.last_hidden_state, self.pooler_output)[i]


class FakeVAE(torch.nn.Module):
    """decode(z).sample -> a fixed-size image that depends on z (images are outside the parity: the reference loops
    end in vae.decode, models/pipelines.py:233,461,591)"""

    def __init__(self):
        super().__init__()
        self._dummy = torch.nn.Parameter(torch.zeros(1), requires_grad=False)

    def decode(self, z):
        img = torch.tanh(z[:, :3].float())
        return types.SimpleNamespace(sample=torch.nn.functional.interpolate(img, size=(16, 16), mode="nearest"))


def model_dict_fakes(ctx_dim=768):
    tok = WordTokenizer()
    return dict(tokenizer=tok, text_encoder=FakeTextEncoder(tok, ctx_dim), vae=FakeVAE())
