"""Drop-in for ``ladiff.models.architectures.mld_clip.MldTextEncoder`` (reference file lines 13-90).

CLIP runs once per prompt in plain torch on the GPU; it is *not* part of the optimised path (north_star: "timed and
reported separately, not optimised").  Two cheap, exact savings over the reference: identical strings in one call are
encoded once (the reference re-encodes B copies of ``""`` for the unconditional half, ladiff.py:259-265) and the
``""`` embedding is cached across calls.

No CLIP weights exist offline, so ``modelpath="synthetic://clip-vit-large-patch14"`` builds the ViT-L/14 *text tower*
architecture (hidden 768, 12 layers, 12 heads, MLP 3072, vocab 49408, 77 positions, projection 768) with random
weights and a deterministic stand-in tokenizer.  A real directory path loads tokenizer + weights like the reference.
"""
from __future__ import annotations

import os
import zlib
from typing import List

import torch
from torch import nn

SYNTHETIC = "synthetic://clip-vit-large-patch14"
BOS, EOS, VOCAB, CTX = 49406, 49407, 49408, 77


class _SyntheticTokenizer:
    """Deterministic word-hash tokenizer with CLIP's framing (BOS, tokens, EOS, EOS padding to 77)."""
    model_max_length = CTX

    def __call__(self, texts, padding="max_length", truncation=True, max_length=CTX, return_tensors="pt"):
        rows = []
        for t in texts:
            ids = [BOS] + [1 + zlib.crc32(w.encode()) % (BOS - 1) for w in t.lower().split()][: max_length - 2] + [EOS]
            rows.append(ids + [EOS] * (max_length - len(ids)))
        return type("Enc", (), {"input_ids": torch.tensor(rows, dtype=torch.long)})()


class MldTextEncoder(nn.Module):

    def __init__(self, modelpath: str, finetune: bool = False, last_hidden_state: bool = False,
                 latent_dim: list = [1, 256], seed: int = 1234) -> None:
        super().__init__()
        self.latent_dim = latent_dim
        if last_hidden_state:
            raise NotImplementedError("last_hidden_state=True (77 text tokens) is not the shipped configuration "
                                      "(configs/modules/text_encoder.yaml:6); the denoiser path assumes one pooled token")
        if modelpath == SYNTHETIC or modelpath is None:
            from transformers import CLIPTextConfig, CLIPTextModelWithProjection
            cfg = CLIPTextConfig(vocab_size=VOCAB, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                                 num_attention_heads=12, max_position_embeddings=CTX, projection_dim=768,
                                 hidden_act="quick_gelu", bos_token_id=BOS, eos_token_id=EOS, pad_token_id=EOS)
            with torch.random.fork_rng(devices=[]):
                torch.manual_seed(seed)
                self.text_model = CLIPTextModelWithProjection(cfg)
            self.tokenizer = _SyntheticTokenizer()
            self._projected = True
            self.text_encoded_dim = 768
        elif os.path.isdir(modelpath):
            from transformers import AutoModel, AutoTokenizer
            if "clip" not in modelpath:
                raise ValueError(f"Model {modelpath} not supported")
            self.tokenizer = AutoTokenizer.from_pretrained(modelpath)
            self.text_model = AutoModel.from_pretrained(modelpath)
            self._projected = False
            self.text_encoded_dim = self.text_model.config.text_config.hidden_size
        else:
            raise FileNotFoundError(f"CLIP path {modelpath!r} does not exist (use {SYNTHETIC!r} for random-init weights)")
        self.name = "clip"
        self.max_length = self.tokenizer.model_max_length
        if not finetune:
            self.text_model.training = False
            for p in self.text_model.parameters():
                p.requires_grad = False
        self.finetune = bool(finetune)
        self._uncond = None
        # the cached "" embedding is a function of the weights: drop it whenever they are (re)loaded
        self.register_load_state_dict_post_hook(lambda module, incompatible: setattr(module, "_uncond", None))

    def _cache_ok(self) -> bool:
        """The reference re-encodes "" on every call (ladiff.py:259-265); the cache is exact only while the weights cannot
        move: never with finetune=True or in training mode."""
        return not self.finetune and not self.training

    def _apply(self, fn, *a, **k):
        self._uncond = None
        return super()._apply(fn, *a, **k)

    @torch.no_grad()
    def _encode(self, texts: List[str]) -> torch.Tensor:
        ids = self.tokenizer(texts, padding="max_length", truncation=True, max_length=self.max_length,
                             return_tensors="pt").input_ids[:, : self.max_length]
        dev = next(self.text_model.parameters()).device
        if self._projected:
            return self.text_model(input_ids=ids.to(dev)).text_embeds
        out = self.text_model.get_text_features(ids.to(dev))
        return out if torch.is_tensor(out) else out.pooler_output   # transformers >= 5 returns a ModelOutput

    def forward(self, texts: List[str]) -> torch.Tensor:
        """List[str] (length n) -> [n, 1, 768]  (reference :50-90, 'clip' branch :75-78)."""
        if not self._cache_ok():
            self._uncond = None
        uniq, index = [], {}
        for t in texts:
            if t not in index and not (t == "" and self._uncond is not None):
                index[t] = len(uniq)
                uniq.append(t)
        emb = self._encode(uniq) if uniq else None
        if "" in index and self._cache_ok():
            self._uncond = emb[index[""]].clone()
        rows = [self._uncond if (t == "" and t not in index) else emb[index[t]] for t in texts]
        return torch.stack(rows, 0).unsqueeze(1)
