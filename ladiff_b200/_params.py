"""Parameter containers with the reference's ``state_dict`` key layout.

The CUDA engine owns the arithmetic; these ``nn.Module`` trees only *hold* parameters under exactly the names the
reference modules give them (SURVEY.md 8b "Weights"), so ``torch.load(ckpt)["state_dict"]`` -> ``load_state_dict(strict=True)``
works unchanged (``demo.py:138-159``) and random initialisation follows the same families (xavier-uniform on every
matrix of the skip transformer: ``operator/cross_attention.py:37-40,108-111``; U(0,1) learned positions:
``operator/position_encoding.py:150-151``).  Stock torch layers are used as holders because their parameter names
(``in_proj_weight``, ``out_proj.weight``, ``emb_layers.1.weight`` ...) are the reference's.
"""
from __future__ import annotations

import torch
from torch import nn


class LearnedPE1D(nn.Module):
    """key: ``<name>.pe`` [500, 1, d]  (operator/position_encoding.py:138-160)"""

    def __init__(self, d_model: int, max_len: int = 500):
        super().__init__()
        self.pe = nn.Parameter(torch.zeros(max_len, 1, d_model))
        nn.init.uniform_(self.pe)


class StylizationParams(nn.Module):
    """architectures/mdiff_transformer.py:137-150"""

    def __init__(self, d: int, time_dim: int, dropout: float):
        super().__init__()
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(time_dim, 2 * d))
        self.norm = nn.LayerNorm(d)
        self.out_layers = nn.Sequential(nn.SiLU(), nn.Dropout(p=dropout), nn.Linear(d, d))


class PostNormEncoderLayerParams(nn.Module):
    """architectures/mdiff_transformer.py:26-47 == operator/cross_attention.py:264-286.  Callable in torch
    (post-norm) because ``LADiffVae.encode`` -- outside the CUDA hot path -- delegates to it."""

    def __init__(self, d: int, nhead: int, ff: int, dropout: float, activation: str = "relu"):
        super().__init__()
        self.d_model = d
        self.self_attn = nn.MultiheadAttention(d, nhead, dropout=dropout)
        self.linear1 = nn.Linear(d, ff)
        self.linear2 = nn.Linear(ff, d)
        self.norm1 = nn.LayerNorm(d)
        self.norm2 = nn.LayerNorm(d)
        self.activation = {"relu": torch.relu, "gelu": nn.functional.gelu}[activation]

    def forward(self, src, src_key_padding_mask=None):
        src2 = self.self_attn(src, src, value=src, key_padding_mask=src_key_padding_mask)[0]
        src = self.norm1(src + src2)
        src2 = self.linear2(self.activation(self.linear1(src)))
        return self.norm2(src + src2)


class CrossAttnParams(nn.Module):
    """architectures/mdiff_transformer.py:206-217 (query/key/norm are held for checkpoint compatibility only:
    with a single text token they are mathematically dead)."""

    def __init__(self, d: int, text_d: int, time_d: int, dropout: float):
        super().__init__()
        self.norm = nn.LayerNorm(d)
        self.text_norm = nn.LayerNorm(text_d)
        self.query = nn.Linear(d, d)
        self.key = nn.Linear(text_d, d)
        self.value = nn.Linear(text_d, d)
        self.proj_out = StylizationParams(d, time_d, dropout)


class FFNParams(nn.Module):
    """architectures/mdiff_transformer.py:249-257"""

    def __init__(self, d: int, ff: int, time_d: int, dropout: float):
        super().__init__()
        self.linear1 = nn.Linear(d, ff)
        self.linear2 = nn.Linear(ff, d)
        self.proj_out = StylizationParams(d, time_d, dropout)


class MDLayerParams(nn.Module):
    """architectures/mdiff_transformer.py:265-291: sa_block is hard-wired to ff=1024 / relu (:287-288)."""

    def __init__(self, d: int, text_d: int, time_d: int, ffn_dim: int, nhead: int, dropout: float):
        super().__init__()
        self.d_model = d
        self.ca_block = CrossAttnParams(d, text_d, time_d, dropout)
        self.ffn = FFNParams(d, ffn_dim, time_d, dropout)
        self.sa_block = PostNormEncoderLayerParams(d, nhead, 1024, dropout, "relu")


class DecoderLayerParams(nn.Module):
    """operator/cross_attention.py:332-353"""

    def __init__(self, d: int, nhead: int, ff: int, dropout: float):
        super().__init__()
        self.d_model = d
        self.self_attn = nn.MultiheadAttention(d, nhead, dropout=dropout)
        self.multihead_attn = nn.MultiheadAttention(d, nhead, dropout=dropout)
        self.linear1 = nn.Linear(d, ff)
        self.linear2 = nn.Linear(ff, d)
        self.norm1 = nn.LayerNorm(d)
        self.norm2 = nn.LayerNorm(d)
        self.norm3 = nn.LayerNorm(d)


class SkipStack(nn.Module):
    """U-Net wired stack: keys input_blocks.i / middle_block / output_blocks.i / linear_blocks.i / norm
    (operator/cross_attention.py:19-40, 90-111)."""

    def __init__(self, make_layer, num_layers: int, d: int):
        super().__init__()
        if num_layers % 2 != 1:
            raise AssertionError("num_layers must be odd")  # reference: `assert num_layers % 2 == 1`
        nb = (num_layers - 1) // 2
        self.d_model = d
        self.num_layers = num_layers
        self.input_blocks = nn.ModuleList([make_layer() for _ in range(nb)])
        self.middle_block = make_layer()
        self.output_blocks = nn.ModuleList([make_layer() for _ in range(nb)])
        self.linear_blocks = nn.ModuleList([nn.Linear(2 * d, d) for _ in range(nb)])
        self.norm = nn.LayerNorm(d)
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def forward_encoder(self, x, src_key_padding_mask=None):
        """non-MD branch (cross_attention.py:48-67); torch path used by ``LADiffVae.encode`` only."""
        xs = []
        for m in self.input_blocks:
            x = m(x, src_key_padding_mask=src_key_padding_mask)
            xs.append(x)
        x = self.middle_block(x, src_key_padding_mask=src_key_padding_mask)
        for m, lin in zip(self.output_blocks, self.linear_blocks):
            x = lin(torch.cat([x, xs.pop()], dim=-1))
            x = m(x, src_key_padding_mask=src_key_padding_mask)
        return self.norm(x)


class EngineBound(nn.Module):
    """Mixin: lazily creates / shares a CUDA ``Engine`` and keeps its packed weights in sync with the parameters."""

    _prefix = ""
    _which = 0

    def _init_engine_state(self, precision: str = "bf16x3", engine_kwargs=None):
        from ._lib import MODES
        if precision not in MODES:
            raise ValueError(f"precision must be one of {sorted(MODES)}, got {precision!r}")
        self.precision = precision
        self._engine = None
        self._engine_kwargs = dict(engine_kwargs or {})
        self._dirty = True
        self.register_load_state_dict_post_hook(lambda module, incompatible: setattr(module, "_dirty", True))

    def _apply(self, fn, *a, **k):
        self._dirty = True
        return super()._apply(fn, *a, **k)

    def bind_engine(self, engine):
        """Share one engine (one weight store, one stream of work) between the denoiser and the VAE of a model."""
        self._engine = engine
        self._dirty = True
        return self

    @property
    def mode(self) -> int:
        from ._lib import MODES
        return MODES[self.precision]

    def mark_dirty(self):
        """Force a re-pack of the CUDA engine's weights at the next call (after out-of-band parameter surgery)."""
        self._dirty = True

    def _param_signature(self):
        """(storage address, in-place version counter) of every parameter: ``p.data.copy_()``, an optimizer step or EMA
        bump ``_version``; ``.to()`` / re-assignment change ``data_ptr`` -- a handful of integer reads per call."""
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def engine(self):
        from ._lib import Engine
        if self._engine is None:
            self._engine = Engine(**self._engine_kwargs)
        sig = self._param_signature()
        if self._dirty or sig != getattr(self, "_packed_sig", None):
            self._engine.set_weights(self.state_dict(), self._prefix)
            self._engine.finalize(self._which)
            self._dirty = False
            self._packed_sig = sig
        return self._engine
