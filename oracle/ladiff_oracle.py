"""CPU oracle for the LADiff sampling hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain torch-CPU restatement of the reference algorithm.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it; the product package
(``ladiff_b200``) never does.

Every function cites the reference file:line (relative to
``/root/reference/src/ladiff``) that it restates.  The restatement is kept
*un-hoisted*: it pads every sequence to ``MAX_IT`` latent rows / ``max(L)``
frames, builds the 7-token self-attention input, recomputes time / text
projections every step, etc. -- exactly the work the reference executes --
so it is both the parity checker for the hoisted CUDA path and the timed CPU
baseline.

Pinning: the reference ships no tests or golden vectors.
The restatement is pinned against the reference's *own modules*, imported
unmodified in the authoring container by ``oracle/make_golden.py``; outputs are
committed under ``tests/golden/`` and re-checked on every run by
``tests/test_oracle.py``.  The DDIM scheduler is the exception: it lives in the
third-party ``diffusers`` package (unpinned in ``src/requirements.txt:23``, not
vendored, not installable here), so ``ddim_*`` below restates the published
DDIM update (Song et al. 2021, eq. 12, epsilon-prediction, "leading" timestep
spacing with ``steps_offset``) anchored on the reference call sites
``models/modeltype/ladiff.py:407-417,491-492`` and the parameters in
``configs/modules/scheduler.yaml:1-14``.  **DDIM scheduler: parity unpinned.**
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

D_MODEL = 256
N_HEAD = 4
MAX_IT = 5            # configs/config_ladiff_humanml3d.yaml:58
FRAME_PER_LATENT = 48  # configs/config_ladiff_humanml3d.yaml:59
TEXT_DIM = 768
NUM_LAYERS = 9


# --------------------------------------------------------------------------
# synthetic weights with the reference's state_dict key layout
# --------------------------------------------------------------------------
def _block_names(num_layers: int = NUM_LAYERS) -> List[str]:
    nb = (num_layers - 1) // 2
    return ([f"input_blocks.{i}" for i in range(nb)] + ["middle_block"]
            + [f"output_blocks.{i}" for i in range(nb)])


def state_dict_spec(nfeats: int = 263, num_layers: int = NUM_LAYERS) -> Dict[str, tuple]:
    """Key -> shape of ``{"denoiser.*", "vae.*"}`` exactly as the reference
    modules name them (architectures/ladiff_denoiser.py:16-151,
    architectures/ladiff_vae.py:33-123, operator/cross_attention.py:19-40,90-111,
    264-353, architectures/mdiff_transformer.py:26-47,137-150,206-217,249-291)."""
    D, FF = D_MODEL, 1024
    spec: Dict[str, tuple] = {}

    def lin(p, o, i):
        spec[p + ".weight"] = (o, i)
        spec[p + ".bias"] = (o,)

    def ln(p):
        spec[p + ".weight"] = (D,)
        spec[p + ".bias"] = (D,)

    def mha(p):
        spec[p + ".in_proj_weight"] = (3 * D, D)
        spec[p + ".in_proj_bias"] = (3 * D,)
        lin(p + ".out_proj", D, D)

    def styl(p):
        lin(p + ".emb_layers.1", 2 * D, D)
        ln(p + ".norm")
        lin(p + ".out_layers.2", D, D)

    # ---- denoiser
    lin("denoiser.time_embedding.linear_1", D, TEXT_DIM)
    lin("denoiser.time_embedding.linear_2", D, D)
    lin("denoiser.emb_proj.1", D, TEXT_DIM)
    spec["denoiser.query_pos.pe"] = (500, 1, D)
    spec["denoiser.mem_pos.pe"] = (500, 1, D)
    for b in _block_names(num_layers):
        p = f"denoiser.encoder.{b}"
        ln(p + ".ca_block.norm")
        ln(p + ".ca_block.text_norm")
        lin(p + ".ca_block.query", D, D)
        lin(p + ".ca_block.key", D, D)
        lin(p + ".ca_block.value", D, D)
        styl(p + ".ca_block.proj_out")
        lin(p + ".ffn.linear1", FF, D)
        lin(p + ".ffn.linear2", D, FF)
        styl(p + ".ffn.proj_out")
        mha(p + ".sa_block.self_attn")
        lin(p + ".sa_block.linear1", 1024, D)   # hard-coded 1024: mdiff_transformer.py:287
        lin(p + ".sa_block.linear2", D, 1024)
        ln(p + ".sa_block.norm1")
        ln(p + ".sa_block.norm2")
    for i in range((num_layers - 1) // 2):
        lin(f"denoiser.encoder.linear_blocks.{i}", D, 2 * D)
    ln("denoiser.encoder.norm")

    # ---- vae
    spec["vae.global_motion_token"] = (2 * MAX_IT, D)
    spec["vae.query_pos_encoder.pe"] = (500, 1, D)
    spec["vae.query_pos_decoder.pe"] = (500, 1, D)
    for b in _block_names(num_layers):
        p = f"vae.encoder.{b}"
        mha(p + ".self_attn")
        lin(p + ".linear1", FF, D)
        lin(p + ".linear2", D, FF)
        ln(p + ".norm1")
        ln(p + ".norm2")
    for i in range((num_layers - 1) // 2):
        lin(f"vae.encoder.linear_blocks.{i}", D, 2 * D)
    ln("vae.encoder.norm")
    for b in _block_names(num_layers):
        p = f"vae.decoder.{b}"
        mha(p + ".self_attn")
        mha(p + ".multihead_attn")
        lin(p + ".linear1", FF, D)
        lin(p + ".linear2", D, FF)
        ln(p + ".norm1")
        ln(p + ".norm2")
        ln(p + ".norm3")
    for i in range((num_layers - 1) // 2):
        lin(f"vae.decoder.linear_blocks.{i}", D, 2 * D)
    ln("vae.decoder.norm")
    lin("vae.skel_embedding", D, nfeats)
    lin("vae.final_layer", nfeats, D)
    return spec


def make_state_dict(seed: int = 1234, nfeats: int = 263, perturb: bool = True) -> SD:
    """Deterministic synthetic weights in the reference's key layout.

    Families follow the reference initialisers: xavier-uniform on every
    dim>1 parameter (operator/cross_attention.py:37-40,108-111 re-initialise
    even the ``zero_module``-d ones), U(0,1) learned PE
    (operator/position_encoding.py:150-151), ``nn.Linear`` default bias range.
    ``perturb=True`` additionally randomises the parameters the reference
    leaves at 0/1 (LayerNorm affine, MHA biases, zero-module biases) so that a
    kernel that drops one of them fails parity instead of passing by luck.
    """
    g = torch.Generator().manual_seed(seed)
    sd: SD = {}
    for k, shp in state_dict_spec(nfeats).items():
        if k.endswith(".pe"):
            v = torch.rand(shp, generator=g)
        elif k.endswith("global_motion_token"):
            v = torch.randn(shp, generator=g)
        elif len(shp) > 1:
            fan_out, fan_in = shp
            a = math.sqrt(6.0 / (fan_in + fan_out))
            v = (torch.rand(shp, generator=g) * 2 - 1) * a
        else:
            is_ln = ".norm" in k or k.endswith("text_norm.weight") or k.endswith("text_norm.bias")
            if is_ln:
                if k.endswith("weight"):
                    v = torch.ones(shp)
                    if perturb:
                        v = v + 0.1 * torch.randn(shp, generator=g)
                else:
                    v = torch.zeros(shp)
                    if perturb:
                        v = 0.1 * torch.randn(shp, generator=g)
            else:
                zero_in_ref = (k.endswith("in_proj_bias") or k.endswith("out_proj.bias")
                               or k.endswith("out_layers.2.bias") or k.endswith("ffn.linear2.bias"))
                if zero_in_ref and not perturb:
                    v = torch.zeros(shp)
                else:
                    v = (torch.rand(shp, generator=g) * 2 - 1) * 0.04
        sd[k] = v.float().contiguous()
    return sd


def sub(sd: SD, prefix: str) -> SD:
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


# --------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------
def lengths_to_mask(lengths: Sequence[int], max_len: Optional[int] = None) -> Tensor:
    """utils/temos_utils.py:10-17"""
    lengths = torch.as_tensor(list(lengths))
    max_len = max_len if max_len else int(lengths.max())
    return torch.arange(max_len).expand(len(lengths), max_len) < lengths.unsqueeze(1)


def max_iter_elements_of(lengths: Sequence[int]) -> Tensor:
    """ceil(L / 48) -- models/modeltype/ladiff.py:379, architectures/ladiff_vae.py:292"""
    return torch.ceil(torch.tensor(list(lengths)) / FRAME_PER_LATENT).to(torch.long)


def latent_mask_of(mie: Tensor, max_iter: int = MAX_IT) -> Tensor:
    """architectures/ladiff_denoiser.py:164-171, architectures/ladiff_vae.py:152-159"""
    m = torch.ones((len(mie), max_iter), dtype=torch.bool)
    for i, e in enumerate(mie):
        m[i, int(e):] = False
    return m


def timestep_embedding(timesteps: Tensor, dim: int = TEXT_DIM, flip_sin_to_cos: bool = True,
                       freq_shift: float = 0.0) -> Tensor:
    """architectures/tools/embeddings.py:245-285 (scale=1, max_period=1e4)."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32)
    exponent = exponent / (half - freq_shift)
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


def multi_head_attention(query: Tensor, key: Tensor, value: Tensor, P: SD, prefix: str,
                         key_padding_mask: Optional[Tensor]) -> Tensor:
    """``nn.MultiheadAttention`` (seq-first, 4 heads, eval) as called at
    operator/cross_attention.py:299-300,368-369,373-376 and
    architectures/mdiff_transformer.py:60-61: packed in-projection, q scaled by
    1/sqrt(64) before q.k^T, -inf on padded keys, softmax, .v, out-projection."""
    Lq, B, E = query.shape
    Lk = key.shape[0]
    H, hd = N_HEAD, E // N_HEAD
    w, b = P[prefix + ".in_proj_weight"], P[prefix + ".in_proj_bias"]
    q = F.linear(query, w[:E], b[:E])
    k = F.linear(key, w[E:2 * E], b[E:2 * E])
    v = F.linear(value, w[2 * E:], b[2 * E:])
    q = q.reshape(Lq, B * H, hd).transpose(0, 1) * (1.0 / math.sqrt(hd))
    k = k.reshape(Lk, B * H, hd).transpose(0, 1)
    v = v.reshape(Lk, B * H, hd).transpose(0, 1)
    att = torch.bmm(q, k.transpose(1, 2))                      # [B*H, Lq, Lk]
    if key_padding_mask is not None:
        neg = torch.zeros(B, 1, 1, Lk).masked_fill(key_padding_mask.view(B, 1, 1, Lk), float("-inf"))
        att = (att.view(B, H, Lq, Lk) + neg).view(B * H, Lq, Lk)
    att = torch.softmax(att, dim=-1)
    out = torch.bmm(att, v).transpose(0, 1).reshape(Lq, B, E)
    return F.linear(out, P[prefix + ".out_proj.weight"], P[prefix + ".out_proj.bias"])


def _ln(x: Tensor, P: SD, prefix: str) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), P[prefix + ".weight"], P[prefix + ".bias"], 1e-5)


def encoder_layer_post(src: Tensor, P: SD, p: str, key_padding_mask, activation) -> Tensor:
    """Post-norm TransformerEncoderLayer: architectures/mdiff_transformer.py:54-67
    (== operator/cross_attention.py:293-307)."""
    src2 = multi_head_attention(src, src, src, P, p + ".self_attn", key_padding_mask)
    src = _ln(src + src2, P, p + ".norm1")
    src2 = F.linear(activation(F.linear(src, P[p + ".linear1.weight"], P[p + ".linear1.bias"])),
                    P[p + ".linear2.weight"], P[p + ".linear2.bias"])
    return _ln(src + src2, P, p + ".norm2")


def stylization(h: Tensor, emb: Tensor, P: SD, p: str) -> Tensor:
    """StylizationBlock.forward: architectures/mdiff_transformer.py:152-163 (h [B,T,D], emb [B,D])."""
    emb_out = F.linear(F.silu(emb), P[p + ".emb_layers.1.weight"], P[p + ".emb_layers.1.bias"]).unsqueeze(1)
    scale, shift = torch.chunk(emb_out, 2, dim=2)
    h = _ln(h, P, p + ".norm") * (1 + scale) + shift
    return F.linear(F.silu(h), P[p + ".out_layers.2.weight"], P[p + ".out_layers.2.bias"])


def linear_cross_attention(x: Tensor, xf: Tensor, emb: Tensor, P: SD, p: str,
                           src_key_padding_mask: Optional[Tensor]) -> Tensor:
    """LinearTemporalCrossAttention.forward: architectures/mdiff_transformer.py:219-247."""
    B, T, D = x.shape
    N = xf.shape[1]
    H = N_HEAD
    if src_key_padding_mask is not None:
        keep = (~src_key_padding_mask).long().unsqueeze(2).repeat(1, 1, D)
    query = F.linear(_ln(x, P, p + ".norm"), P[p + ".query.weight"], P[p + ".query.bias"])
    key = F.linear(_ln(xf, P, p + ".text_norm"), P[p + ".key.weight"], P[p + ".key.bias"])
    query = F.softmax(query.view(B, T, H, -1), dim=-1)
    key = F.softmax(key.view(B, N, H, -1), dim=1)
    value = F.linear(_ln(xf, P, p + ".text_norm"), P[p + ".value.weight"], P[p + ".value.bias"]).view(B, N, H, -1)
    attention = torch.einsum("bnhd,bnhl->bhdl", key, value)
    if src_key_padding_mask is not None:
        query = query * keep.view(B, T, H, -1)
    y = torch.einsum("bnhd,bhdl->bnhl", query, attention).reshape(B, T, D)
    return x + stylization(y, emb, P, p + ".proj_out")


def ffn_block(x: Tensor, emb: Tensor, P: SD, p: str) -> Tensor:
    """FFN.forward: architectures/mdiff_transformer.py:259-262 (exact erf GELU)."""
    y = F.linear(F.gelu(F.linear(x, P[p + ".linear1.weight"], P[p + ".linear1.bias"])),
                 P[p + ".linear2.weight"], P[p + ".linear2.bias"])
    return x + stylization(y, emb, P, p + ".proj_out")


def md_layer(x: Tensor, xf: Tensor, emb: Tensor, P: SD, p: str, src_key_padding_mask: Tensor,
             trace: Optional[dict] = None) -> Tensor:
    """LinearTemporalDiffusionTransformerDecoderLayer.forward:
    architectures/mdiff_transformer.py:294-321.  x [T,B,D], xf [1,B,D], emb [1,B,D]."""
    aug = torch.cat([src_key_padding_mask, torch.zeros((src_key_padding_mask.shape[0], 2), dtype=torch.bool)], dim=1)
    latent_in = x.shape[0]
    seq = torch.cat([x, xf, emb], dim=0)                                   # :311
    seq = encoder_layer_post(seq, P, p + ".sa_block", aug, F.relu)         # :312 ('relu', ff=1024: :287)
    xb = seq[:latent_in].permute(1, 0, 2)                                  # :313
    if trace is not None:
        trace[p + ".sa"] = xb.clone()
    xb = linear_cross_attention(xb, xf.permute(1, 0, 2), emb.permute(1, 0, 2).squeeze(1), P, p + ".ca_block",
                                src_key_padding_mask)                      # :317
    if trace is not None:
        trace[p + ".ca"] = xb.clone()
    xb = ffn_block(xb, emb.permute(1, 0, 2).squeeze(1), P, p + ".ffn")     # :318
    if trace is not None:
        trace[p + ".ffn"] = xb.clone()
    return xb.permute(1, 0, 2)


def skip_encoder_md(src: Tensor, xf: Tensor, emb: Tensor, P: SD, p: str, src_key_padding_mask: Tensor,
                    trace: Optional[dict] = None) -> Tensor:
    """SkipTransformerEncoder.forward, MD_trans branch: operator/cross_attention.py:69-85."""
    nb = (NUM_LAYERS - 1) // 2
    x, xs = src, []
    for i in range(nb):
        x = md_layer(x, xf, emb, P, f"{p}.input_blocks.{i}", src_key_padding_mask, trace)
        xs.append(x)
    x = md_layer(x, xf, emb, P, f"{p}.middle_block", src_key_padding_mask, trace)
    for i in range(nb):
        x = torch.cat([x, xs.pop()], dim=-1)
        x = F.linear(x, P[f"{p}.linear_blocks.{i}.weight"], P[f"{p}.linear_blocks.{i}.bias"])
        x = md_layer(x, xf, emb, P, f"{p}.output_blocks.{i}", src_key_padding_mask, trace)
    return _ln(x, P, p + ".norm")


def denoiser_forward(sd: SD, sample: Tensor, timestep: Tensor, encoder_hidden_states: Tensor,
                     max_iter_elements: Optional[Tensor], trace: Optional[dict] = None,
                     enclat: Optional[Tensor] = None) -> Tensor:
    """LADiffDenoiser.forward, text condition / trans_enc / MD_TRANS:
    architectures/ladiff_denoiser.py:153-295.  sample [2B,T,256], timestep 0-dim,
    encoder_hidden_states [2B,1,768] -> [2B,T,256].  ``enclat`` [2B,k,256] (ARDIFF conditioning latents, :218-219,
    :247-248) is appended to the token sequence; that branch passes no max_iter_elements -> no key-padding mask (:252-255)."""
    T_in = sample.shape[1]
    n_tok = T_in + (enclat.shape[1] if enclat is not None else 0)
    if max_iter_elements is not None:
        latent_mask = latent_mask_of(max_iter_elements, n_tok)                          # :164-171
    else:
        latent_mask = torch.ones((sample.shape[0], n_tok), dtype=torch.bool)
    sample = sample.permute(1, 0, 2)                                                    # :175
    if enclat is not None:
        sample_in = torch.cat((sample, enclat.permute(1, 0, 2)), dim=0)                 # :219,248
    else:
        sample_in = sample
    timesteps = timestep.expand(sample.shape[1]).clone()                                # :184
    time_emb = timestep_embedding(timesteps).to(sample.dtype)                           # :185-186
    time_emb = F.linear(F.silu(F.linear(time_emb, sd["denoiser.time_embedding.linear_1.weight"],
                                        sd["denoiser.time_embedding.linear_1.bias"])),
                        sd["denoiser.time_embedding.linear_2.weight"],
                        sd["denoiser.time_embedding.linear_2.bias"]).unsqueeze(0)       # :188
    text_emb = encoder_hidden_states.permute(1, 0, 2)                                   # :193
    text_emb_latent = F.linear(F.relu(text_emb), sd["denoiser.emb_proj.1.weight"],
                               sd["denoiser.emb_proj.1.bias"])                          # :72-73,198
    xseq = sample_in + sd["denoiser.query_pos.pe"][:sample_in.shape[0]]                 # :251
    tokens = skip_encoder_md(xseq, text_emb_latent, time_emb, sd, "denoiser.encoder", ~latent_mask, trace)
    return tokens[:sample.shape[0]].permute(1, 0, 2)                                    # :272,292


# ---- LA-VAE decoder ---------------------------------------------------------
def decoder_layer_post(tgt: Tensor, memory: Tensor, P: SD, p: str, tgt_kpm: Tensor, mem_kpm: Tensor) -> Tensor:
    """TransformerDecoderLayer.forward_post: operator/cross_attention.py:358-413 (pos/query_pos None)."""
    t2 = multi_head_attention(tgt, tgt, tgt, P, p + ".self_attn", tgt_kpm)
    tgt = _ln(tgt + t2, P, p + ".norm1")
    t2 = multi_head_attention(tgt, memory, memory, P, p + ".multihead_attn", mem_kpm)
    tgt = _ln(tgt + t2, P, p + ".norm2")
    t2 = F.linear(F.gelu(F.linear(tgt, P[p + ".linear1.weight"], P[p + ".linear1.bias"])),
                  P[p + ".linear2.weight"], P[p + ".linear2.bias"])
    return _ln(tgt + t2, P, p + ".norm3")


def skip_decoder(tgt: Tensor, memory: Tensor, P: SD, p: str, tgt_kpm: Tensor, mem_kpm: Tensor) -> Tensor:
    """SkipTransformerDecoder.forward: operator/cross_attention.py:113-153."""
    nb = (NUM_LAYERS - 1) // 2
    x, xs = tgt, []
    for i in range(nb):
        x = decoder_layer_post(x, memory, P, f"{p}.input_blocks.{i}", tgt_kpm, mem_kpm)
        xs.append(x)
    x = decoder_layer_post(x, memory, P, f"{p}.middle_block", tgt_kpm, mem_kpm)
    for i in range(nb):
        x = torch.cat([x, xs.pop()], dim=-1)
        x = F.linear(x, P[f"{p}.linear_blocks.{i}.weight"], P[f"{p}.linear_blocks.{i}.bias"])
        x = decoder_layer_post(x, memory, P, f"{p}.output_blocks.{i}", tgt_kpm, mem_kpm)
    return _ln(x, P, p + ".norm")


def vae_decode(sd: SD, z: Tensor, lengths: Sequence[int]) -> Tensor:
    """LADiffVae.decode (arch encoder_decoder, pe_type mld): architectures/ladiff_vae.py:288-362.
    z [T,B,256] -> feats [B, max(lengths), nfeats], padded frames exactly zero."""
    mask = lengths_to_mask(lengths)                                                     # :290
    latent_mask = latent_mask_of(max_iter_elements_of(lengths), z.shape[0])             # :292-295
    bs, nframes = mask.shape
    queries = torch.zeros(nframes, bs, D_MODEL) + sd["vae.query_pos_decoder.pe"][:nframes]   # :299,334
    out = skip_decoder(queries, z, sd, "vae.decoder", ~mask, ~latent_mask)              # :337-345
    out = F.linear(out, sd["vae.final_layer.weight"], sd["vae.final_layer.bias"])       # :356
    out[~mask.T] = 0                                                                    # :358
    return out.permute(1, 0, 2)                                                         # :360


# ---- DDIM (third-party diffusers; parity unpinned, see module docstring) ----
def ddim_alphas_cumprod(num_train_timesteps: int = 1000, beta_start: float = 0.00085,
                        beta_end: float = 0.012) -> Tensor:
    """'scaled_linear' betas, configs/modules/scheduler.yaml:6-9."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def ddim_timesteps(n: int, num_train_timesteps: int = 1000, steps_offset: int = 1) -> np.ndarray:
    """'leading' spacing + steps_offset (scheduler.yaml:14): n=50 -> [981, 961, ..., 21, 1]."""
    ratio = num_train_timesteps // n
    return (np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64) + steps_offset


def ddim_step(eps: Tensor, t: int, sample: Tensor, acp: Tensor, n: int, num_train_timesteps: int = 1000) -> Tensor:
    """eta=0, epsilon prediction, clip_sample false, set_alpha_to_one false
    (scheduler.yaml:3,10-13).  Call site models/modeltype/ladiff.py:491-492."""
    prev_t = t - num_train_timesteps // n
    a_t = acp[t]
    a_p = acp[prev_t] if prev_t >= 0 else acp[0]
    x0 = (sample - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
    return a_p ** 0.5 * x0 + (1 - a_p) ** 0.5 * eps


def ddpm_step(eps: Tensor, t: int, sample: Tensor, acp: Tensor, n: int, noise: Optional[Tensor],
              num_train_timesteps: int = 1000) -> Tensor:
    """diffusers DDPMScheduler.step, epsilon prediction, variance_type fixed_small, clip_sample false
    (configs/modules/scheduler.yaml:16-29; Ho et al. 2020 eq. 7); ``noise`` is the variance noise the
    scheduler draws itself.  Third-party: parity unpinned like DDIM."""
    prev_t = t - num_train_timesteps // n
    a_t = acp[t].double()
    a_p = acp[prev_t].double() if prev_t >= 0 else torch.tensor(1.0, dtype=torch.float64)
    cur_alpha = a_t / a_p
    cur_beta = 1 - cur_alpha
    x0 = (sample - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
    prev = (a_p ** 0.5 * cur_beta / (1 - a_t)) * x0 + (cur_alpha ** 0.5 * (1 - a_p) / (1 - a_t)) * sample
    if t > 0:
        var = torch.clamp((1 - a_p) / (1 - a_t) * cur_beta, min=1e-20)
        prev = prev + var ** 0.5 * noise
    return prev.to(sample.dtype)


def ddpm_timesteps(n: int, num_train_timesteps: int = 1000) -> np.ndarray:
    return np.arange(0, num_train_timesteps, num_train_timesteps // n)[::-1].copy().astype(np.int64)


# ---- the sampling loop -----------------------------------------------------
def initial_latents(noise: Tensor, lengths: Sequence[int]) -> Tensor:
    """models/modeltype/ladiff.py:379-390,407: randn [B,5,256] (injected), rows >= m_i zeroed, x init_noise_sigma(=1)."""
    mie = max_iter_elements_of(lengths)
    lat = noise.clone()
    for i, e in enumerate(mie):
        lat[i, int(e):] = 0
    return lat * 1.0


def diffusion_reverse(sd: SD, encoder_hidden_states: Tensor, lengths: Sequence[int], noise: Tensor,
                      num_inference_steps: int = 50, guidance_scale: float = 7.5,
                      record: Optional[dict] = None, denoiser_fn=None, scheduler: str = "ddim",
                      step_noise: Optional[Tensor] = None) -> Tensor:
    """LADIFF._diffusion_reverse, IDEA 'ard' / ARDIFF False / LAD branch:
    models/modeltype/ladiff.py:333-340,378-390,406-417,470-500,562-566.
    encoder_hidden_states [2B,1,768] (uncond rows first), noise [B,5,256] -> latents [5,B,256]."""
    mie = max_iter_elements_of(lengths)
    latents = initial_latents(noise, lengths)
    acp = ddim_alphas_cumprod()
    ts = ddim_timesteps(num_inference_steps) if scheduler == "ddim" else ddpm_timesteps(num_inference_steps)
    mie2 = torch.cat([mie] * 2)
    for i, t in enumerate(ts):
        x2 = torch.cat([latents] * 2)                                                   # :472-474
        if denoiser_fn is None:
            noise_pred = denoiser_forward(sd, x2, torch.tensor(int(t)), encoder_hidden_states, mie2)
        else:   # e.g. the reference's own LADiffDenoiser (oracle/make_golden.py)
            noise_pred = denoiser_fn(x2, torch.tensor(int(t)), encoder_hidden_states, mie2)
        u, c = noise_pred.chunk(2)                                                      # :488
        noise_pred = u + guidance_scale * (c - u)                                       # :489-490
        if scheduler == "ddim":
            latents = ddim_step(noise_pred, int(t), latents, acp, num_inference_steps)  # :491-492
        else:       # step_noise [n,B,5,256]: what DDPMScheduler.step would draw
            latents = ddpm_step(noise_pred, int(t), latents, acp, num_inference_steps, step_noise[i])
        if record is not None and (i + 1) in record.get("steps", ()):
            record[f"latents_after_{i + 1}"] = latents.clone()
    latents = latents.permute(1, 0, 2)                                                  # :500
    for i, e in enumerate(mie):
        latents[int(e):, i] = 0                                                         # :564-566
    return latents


def diffusion_reverse_ardiff(sd: SD, encoder_hidden_states: Tensor, lengths: Sequence[int], noise: Tensor,
                             num_inference_steps: int = 50, guidance_scale: float = 7.5,
                             motion_conditioning: str = "last", denoiser_fn=None) -> Tensor:
    """LADIFF._diffusion_reverse, ARDIFF branch: models/modeltype/ladiff.py:343-365,419-467,562-570.
    noise [B, ar_iterations, 256] -> latents [MAX_IT, B, 256]."""
    ar_iterations = -(-max(lengths) // FRAME_PER_LATENT)                                # :349-356
    acp = ddim_alphas_cumprod()
    ts = ddim_timesteps(num_inference_steps)
    final = None
    for k in range(ar_iterations):
        lat = noise[:, k:k + 1].clone()                                                 # :423
        if k > 0:
            enclat = final[:, :k] if motion_conditioning in ("full", "middle") else final[:, k - 1:k]   # :425-431
            enclat = torch.cat([enclat] * 2)
        else:
            enclat = None
        for t in ts:
            x2 = torch.cat([lat] * 2)
            if denoiser_fn is None:
                pred = denoiser_forward(sd, x2, torch.tensor(int(t)), encoder_hidden_states, None, enclat=enclat)
            else:
                pred = denoiser_fn(x2, torch.tensor(int(t)), encoder_hidden_states, enclat)
            u, c = pred.chunk(2)
            lat = ddim_step(u + guidance_scale * (c - u), int(t), lat, acp, num_inference_steps)
        final = lat if final is None else torch.cat((final, lat), dim=1)                # :462
    latents = final.permute(1, 0, 2)                                                    # :466
    for i, e in enumerate(max_iter_elements_of(lengths)):
        latents[int(e):, i] = 0                                                         # :564-566
    if latents.shape[0] < MAX_IT:                                                       # :567-569
        latents = torch.cat((latents, torch.zeros((MAX_IT - latents.shape[0],) + tuple(latents.shape[1:]))), dim=0)
    return latents


# ---- LA-VAE encoder ("next" row f3) -------------------------------------------------
def skip_encoder_plain(src: Tensor, P: SD, p: str, src_key_padding_mask: Tensor) -> Tensor:
    """SkipTransformerEncoder.forward, non-MD branch: operator/cross_attention.py:48-67 with
    TransformerEncoderLayer.forward_post (:293-307, gelu)."""
    nb = (NUM_LAYERS - 1) // 2
    x, xs = src, []
    for i in range(nb):
        x = encoder_layer_post(x, P, f"{p}.input_blocks.{i}", src_key_padding_mask, F.gelu)
        xs.append(x)
    x = encoder_layer_post(x, P, f"{p}.middle_block", src_key_padding_mask, F.gelu)
    for i in range(nb):
        x = torch.cat([x, xs.pop()], dim=-1)
        x = F.linear(x, P[f"{p}.linear_blocks.{i}.weight"], P[f"{p}.linear_blocks.{i}.bias"])
        x = encoder_layer_post(x, P, f"{p}.output_blocks.{i}", src_key_padding_mask, F.gelu)
    return _ln(x, P, p + ".norm")


def vae_encode(sd: SD, features: Tensor, lengths: Sequence[int], eps: Optional[Tensor] = None):
    """LADiffVae.encode (LAD, JOINT_DISTRO_FIX false, MLP_DIST false): architectures/ladiff_vae.py:162-286.
    features [B, max(lengths), nfeats] -> (latent [5,B,256], mu [5,B,256], std [5,B,256], max_iter_elements).
    ``eps`` [5,B,256] is the standard-normal draw of ``dist.rsample()`` (latent = mu + std * eps)."""
    bs = features.shape[0]
    mask = lengths_to_mask(lengths, features.shape[1])                                  # :178
    x = F.linear(features, sd["vae.skel_embedding.weight"], sd["vae.skel_embedding.bias"]).permute(1, 0, 2)   # :182-186
    dist = torch.tile(sd["vae.global_motion_token"][:, None, :], (1, bs, 1))            # :189
    mie = max_iter_elements_of(lengths)                                                 # :198
    dm = latent_mask_of(mie, MAX_IT)
    aug_mask = torch.cat((dm, dm, mask), 1)                                             # :203-210
    xseq = torch.cat((dist, x), 0)                                                      # :213
    xseq = xseq + sd["vae.query_pos_encoder.pe"][:xseq.shape[0]]                        # :220
    out = skip_encoder_plain(xseq, sd, "vae.encoder", ~aug_mask)[:dist.shape[0]]        # :221-222
    mu, logvar = out[:MAX_IT], out[MAX_IT:]                                             # :258-259
    std = logvar.exp().pow(0.5)                                                         # :262
    latent = mu + std * (eps if eps is not None else torch.randn(mu.shape))             # :263-264
    for i, e in enumerate(mie):
        latent[int(e):, i] = 0                                                          # :265-268
    return latent, mu, std, mie


def sample_motion(sd: SD, encoder_hidden_states: Tensor, lengths: Sequence[int], noise: Tensor,
                  num_inference_steps: int = 50, guidance_scale: float = 7.5) -> Tensor:
    """LADIFF.forward after the text encoder: models/modeltype/ladiff.py:266,283."""
    z = diffusion_reverse(sd, encoder_hidden_states, lengths, noise, num_inference_steps, guidance_scale)
    return vae_decode(sd, z, lengths)


# ---- feats2joints ("next" row f1) -------------------------------------------
def qrot(q: Tensor, v: Tensor) -> Tensor:
    """data/humanml/common/quaternion.py:54-73"""
    qvec = q[..., 1:]
    uv = torch.cross(qvec, v, dim=-1)
    uuv = torch.cross(qvec, uv, dim=-1)
    return v + 2 * (q[..., :1] * uv + uuv)


def recover_from_ric(data: Tensor, joints_num: int) -> Tensor:
    """data/humanml/scripts/motion_process.py:355-381,415-430; q_inv = conj: quaternion.py:16-20."""
    rot_vel = data[..., 0]
    r_rot_ang = torch.zeros_like(rot_vel)
    r_rot_ang[..., 1:] = rot_vel[..., :-1]
    r_rot_ang = torch.cumsum(r_rot_ang, dim=-1)
    r_rot_quat = torch.zeros(data.shape[:-1] + (4,), dtype=data.dtype)
    r_rot_quat[..., 0] = torch.cos(r_rot_ang)
    r_rot_quat[..., 2] = torch.sin(r_rot_ang)
    r_pos = torch.zeros(data.shape[:-1] + (3,), dtype=data.dtype)
    r_pos[..., 1:, [0, 2]] = data[..., :-1, 1:3]
    q_inv = r_rot_quat * torch.tensor([1.0, -1.0, -1.0, -1.0], dtype=data.dtype)
    r_pos = qrot(q_inv, r_pos)
    r_pos = torch.cumsum(r_pos, dim=-2)
    r_pos[..., 1] = data[..., 3]
    positions = data[..., 4:(joints_num - 1) * 3 + 4]
    positions = positions.view(positions.shape[:-1] + (-1, 3))
    positions = qrot(q_inv[..., None, :].expand(positions.shape[:-1] + (4,)), positions)
    positions[..., 0] += r_pos[..., 0:1]
    positions[..., 2] += r_pos[..., 2:3]
    return torch.cat([r_pos.unsqueeze(-2), positions], dim=-2)


def feats2joints(features: Tensor, mean: Tensor, std: Tensor, njoints: int) -> Tensor:
    """data/HumanML3D.py:44-48"""
    return recover_from_ric(features * std + mean, njoints)


# ---- seeded synthetic inputs shared by tests and bench ---
def synthetic_inputs(B: int, seed: int = 1234, ragged: bool = True, fixed_len: int = 196):
    g = torch.Generator().manual_seed(seed)
    text = torch.randn((2 * B, 1, TEXT_DIM), generator=g)
    noise = torch.randn((B, MAX_IT, D_MODEL), generator=g)
    if ragged:
        lengths = (np.random.default_rng(seed).integers(10, 50, size=B) * 4).tolist()
    else:
        lengths = [fixed_len] * B
    return text, noise, [int(x) for x in lengths]
