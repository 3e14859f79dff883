"""What the layout-grounded denoising path needs from its surroundings, as one small interface.

The reference reaches these through module globals (`models.model_dict`: tokenizer, text_encoder, vae, sam_model -
generation/lmd_plus.py:12-19).  They are NOT part of the B200 hot path (CLIP, VAE and SAM are
third-party models marked out of scope / "next"), so the path takes them as callables:

  encode_prompts(prompts, negative_prompt)      -> (uncond [1,T,ctx], cond [len(prompts),T,ctx])   models/models.py:63-89
  phrase_embeddings(phrases)                    -> [len(phrases), 768]  (CLIP pooler_output)        models/pipelines.py:303-304
  phrase_indices(prompt, phrases, words, add_suffix) -> (object_positions, word_token_indices, prompt)
                                                                                                   utils/guidance.py:32-89
  decode(latents [B,4,H,W])                     -> uint8 [B, 8H, 8W, 3] or None                    models/pipelines.py:117-127
  refine_mask(image, box, H, W, token_attn)     -> bool [H, W]                                     models/sam.py:113-213

`SyntheticEnv` is the offline stand-in used by tests and bench.py (no CLIP vocabulary, SD weights or SAM weights exist
in this environment): seeded embeddings, a whitespace tokenizer, the box raster as "SAM" mask, no VAE.
`ReferenceEnv` adapts a reference-style model_dict (real tokenizer / text_encoder / vae / sam) to the same interface.
"""
import hashlib

import numpy as np
import torch

from .latents import box_to_mask


def _seed_of(s):
    return int.from_bytes(hashlib.sha256(s.encode()).digest()[:4], "little")


class SyntheticEnv:
    def __init__(self, ctx_dim=768, T=77, latent_hw=(64, 64), cache_device=None, vae_decoder=None):
        self.ctx_dim, self.T = ctx_dim, T
        self.vae_decoder = vae_decoder      # lgd_b200.vae.B200VAEDecoder (synthetic weights offline) or None
        self.latent_hw = latent_hw
        # cache_device set: embeddings are memoised ON the device (inputs resident in HBM);
        # unset: every call produces pinned host tensors that the path copies host->device itself.
        self.cache_device = cache_device
        self._cache = {}
        self.bytes_out = 0

    # --- tokenisation: <bos> word word ... <eos>, one token per whitespace-separated word
    def tokens(self, prompt):
        return ["<bos>"] + prompt.replace(",", " ,").split() + ["<eos>"]

    def phrase_indices(self, prompt, phrases, words=None, add_suffix=True):
        """same contract as utils/guidance.py:32-89 (suffixing with "| phrase" when absent, last token of the word)"""
        for ph in phrases:
            if ph not in prompt:
                if not add_suffix:
                    # the caller has already encoded `prompt` (per-box generation, generation/lmd.py:76-85): a suffix
                    # added here would shift the token indices away from the encoded text
                    raise ValueError(f"phrase {ph!r} not found in prompt {prompt!r}")
                prompt += "| " + ph
        toks = self.tokens(prompt)
        positions, word_idx = [], []
        for i, ph in enumerate(phrases):
            pt = self.tokens(ph)[1:-1]
            start = next(s for s in range(len(toks)) if toks[s:s + len(pt)] == pt)
            positions.append(list(range(start, start + len(pt))))
            if words is not None:
                wt = self.tokens(words[i])[-2]
                word_idx.append(start + pt.index(wt))
            else:
                word_idx.append(positions[0][-1])
        return positions, word_idx, prompt

    def _embed(self, text, rows):
        key = (text, rows)
        if self.cache_device is not None and key in self._cache:
            return self._cache[key]
        g = torch.Generator().manual_seed(_seed_of(text))
        t = torch.randn(rows, self.ctx_dim, generator=g)
        if self.cache_device is not None:
            t = t.to(self.cache_device)
            self._cache[key] = t
        else:
            if torch.cuda.is_available():
                t = t.pin_memory()
            self.bytes_out += t.numel() * t.element_size()
        return t

    def encode_prompts(self, prompts, negative_prompt=""):
        uncond = self._embed("neg:" + negative_prompt, self.T)[None]
        cond = torch.stack([self._embed("pos:" + p, self.T) for p in prompts])
        return uncond, cond

    def phrase_embeddings(self, phrases):
        return torch.stack([self._embed("phrase:" + p, 1)[0] for p in phrases])

    def decode(self, latents):
        """models/pipelines.py:117-127 on the B200 VAE decoder when one is configured: uint8 [B, 8H, 8W, 3] (host)"""
        if self.vae_decoder is None:
            return None
        return self.vae_decoder.decode(latents).cpu().numpy()

    def refine_mask(self, image, box, H, W, token_attn=None):
        return box_to_mask(box, H, W).bool()


class ReferenceEnv:
    """adapter over the reference's model_dict (tokenizer, text_encoder, vae[, sam]) - see INTEGRATION.md"""

    def __init__(self, model_dict, refine_mask=None, device="cuda", sam_predict=None, sam_kwargs=None):
        """refine_mask: a complete replacement of the SAM step, `f(image, box, token_attn) -> [H, W] mask`;
        sam_predict: only the SAM network, `f(image, input_boxes=None, input_points=None) -> (three masks at image
        resolution, three predicted IoUs)` - prompt construction and candidate selection then run in mask_refine.py
        (`sam_predict_from(model_dict)` wraps the reference's own sam_model / sam_processor); sam_kwargs: overrides of the
        thresholds (mask_refine.refine_attn / refine_box keyword arguments)"""
        self.md = model_dict
        self.device = device
        self._refine = refine_mask
        self._sam_predict = sam_predict
        self._sam_kwargs = dict(sam_kwargs or {})

    def _token_map(self, prompt):
        ids = self.md.tokenizer([prompt], padding="do_not_pad", max_length=77, return_tensors="np")["input_ids"][0]
        return [self.md.tokenizer._convert_id_to_token(int(i)) for i in ids.tolist()]

    def phrase_indices(self, prompt, phrases, words=None, add_suffix=True):
        for ph in phrases:
            if ph not in prompt:
                if not add_suffix:
                    # the caller has already encoded `prompt` (per-box generation, generation/lmd.py:76-85): a suffix
                    # added here would shift the token indices away from the encoded text
                    raise ValueError(f"phrase {ph!r} not found in prompt {prompt!r}")
                prompt += "| " + ph
        joined = " ".join(self._token_map(prompt))
        positions, word_idx = [], []
        for i, ph in enumerate(phrases):
            pt = self._token_map(ph)[1:-1]
            first = len(joined[:joined.index(" ".join(pt)) - 1].split(" "))
            positions.append(list(range(first, first + len(pt))))
            if words is None:
                word_idx.append(positions[0][-1])
            else:
                word_idx.append(first + pt.index(self._token_map(words[i])[-2]))
        return positions, word_idx, prompt

    @torch.no_grad()
    def encode_prompts(self, prompts, negative_prompt=""):
        tok, enc = self.md.tokenizer, self.md.text_encoder
        ti = tok(prompts, padding="max_length", max_length=tok.model_max_length, truncation=True, return_tensors="pt")
        ui = tok([negative_prompt], padding="max_length", max_length=ti.input_ids.shape[-1], return_tensors="pt")
        dev = next(enc.parameters()).device
        return enc(ui.input_ids.to(dev))[0].float().cpu(), enc(ti.input_ids.to(dev))[0].float().cpu()

    @torch.no_grad()
    def phrase_embeddings(self, phrases):
        tok, enc = self.md.tokenizer, self.md.text_encoder
        dev = next(enc.parameters()).device
        inp = tok(phrases, padding=True, return_tensors="pt").to(dev)
        return enc(**inp).pooler_output.float().cpu()

    @torch.no_grad()
    def decode(self, latents):
        vae = self.md.vae
        p = next(vae.parameters())
        img = vae.decode((latents / 0.18215).to(p.device, p.dtype)).sample
        img = (img / 2 + 0.5).clamp(0, 1).float().cpu().permute(0, 2, 3, 1).numpy()
        return (img * 255).round().astype("uint8")

    def refine_mask(self, image, box, H, W, token_attn=None):
        """refine_mask(image, box, token_attn) -> [H, W] mask: the SAM step between the phases (models/sam.py:113-213;
        LMD prompts SAM with points from `token_attn`, LMD+ with the box).  Without a callable: the box raster."""
        if self._refine is not None:
            return torch.as_tensor(self._refine(image, box, token_attn)).bool()
        if self._sam_predict is not None:
            from . import mask_refine as MR
            height, width = int(image.shape[0]), int(image.shape[1])
            if token_attn is not None:      # LMD: models/sam.py sam_refine_attn
                m, _ = MR.refine_attn(self._sam_predict, image, np.asarray(torch.as_tensor(token_attn).cpu()), height, width,
                                      H, W, **self._sam_kwargs)
            else:                           # LMD+: models/sam.py sam_refine_box
                m, _ = MR.refine_box(self._sam_predict, image, box, height, width, H, W, **self._sam_kwargs)
            return torch.as_tensor(m).bool()
        return box_to_mask(box, H, W).bool()


def sam_predict_from(model_dict, device="cuda"):
    """the SAM network call of models/sam.py:25-47 as a `sam_predict` hook over the reference's own
    `model_dict["sam_model"]` / `model_dict["sam_processor"]` (transformers SamModel / SamProcessor - third-party, not
    part of this library): three candidate masks at image resolution + their predicted IoUs for one image / one prompt"""
    model, proc = model_dict["sam_model"], model_dict["sam_processor"]

    @torch.no_grad()
    def predict(image, input_boxes=None, input_points=None):
        with torch.autocast(device):
            inputs = proc(image, input_points=input_points, input_boxes=input_boxes, return_tensors="pt").to(device)
            outputs = model(**inputs)
        masks = proc.image_processor.post_process_masks(outputs.pred_masks.cpu().float(), inputs["original_sizes"].cpu(),
                                                        inputs["reshaped_input_sizes"].cpu())
        return masks[0][0], outputs.iou_scores.cpu().numpy()[0, 0]

    return predict
