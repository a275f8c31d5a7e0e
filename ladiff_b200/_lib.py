"""ctypes binding of ``include/ladiff_b200.h`` and the ``Engine`` object the mirror classes share.

There is no CPU fallback: if the shared library is missing or no sm_100 device is
present every compute path raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LADIFF_LIB") or os.path.join(_HERE, "_C", "libladiff_b200.so")   # LADIFF_LIB: A/B another build

MODE_FP32, MODE_BF16X3, MODE_BF16 = 0, 1, 2
MODES = {"fp32": MODE_FP32, "bf16x3": MODE_BF16X3, "bf16": MODE_BF16}

EPI = {"bias": 0, "relu": 1, "gelu": 2, "res": 3, "ln": 4, "ln_mod_silu": 5, "silu": 6}


class LadiffConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("nfeats", "num_layers", "latent_dim", "num_heads", "ff_size", "text_dim",
                                         "max_it", "frame_per_latent", "max_frames", "use_cuda_graph")]


_lib = None


def load_library() -> C.CDLL:
    """Loads libladiff_b200.so (built by ``python -m ladiff_b200.build``); raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: build it with `python -m ladiff_b200.build` "
                           "(the CUDA extension is mandatory, there is no fallback path)")
    lib = C.CDLL(LIB_PATH)
    missing = [s for s in EXPORTS if not hasattr(lib, s)]
    if missing and not os.environ.get("LADIFF_LIB"):      # an A/B build of an older revision (LADIFF_LIB) may lack newer entry points
        raise RuntimeError(f"{LIB_PATH} does not export {missing}: rebuild with `python -m ladiff_b200.build`")
    for s in missing:
        setattr(lib, s, None)
    vp, i32, f32, i64 = C.c_void_p, C.c_int32, C.c_float, C.c_int64
    pi32, pf32 = C.POINTER(C.c_int32), C.POINTER(C.c_float)
    lib.ladiff_abi_version.restype = C.c_int
    lib.ladiff_create.argtypes = [C.POINTER(LadiffConfig), C.POINTER(vp)]
    lib.ladiff_destroy.argtypes = [vp]
    lib.ladiff_destroy.restype = None
    lib.ladiff_last_error.argtypes = [vp]
    lib.ladiff_last_error.restype = C.c_char_p
    lib.ladiff_set_weight.argtypes = [vp, C.c_char_p, vp, C.POINTER(i64), i32, vp]
    lib.ladiff_finalize_weights.argtypes = [vp, i32, vp]
    lib.ladiff_diffusion_reverse.argtypes = [vp, vp, pi32, i32, vp, i32, pi32, pf32, pf32, f32, i32, vp, vp]
    lib.ladiff_diffusion_reverse_ex.argtypes = [vp, vp, pi32, pi32, i32, vp, i32, pi32, pf32, pf32, pf32, vp, C.c_uint64, i32, f32,
                                                i32, vp, vp]
    lib.ladiff_denoiser_forward.argtypes = [vp, vp, i32, vp, pi32, i32, i32, vp, vp]
    lib.ladiff_cfg_ddim_step.argtypes = [vp, vp, vp, i32, f32, f32, f32, vp]
    lib.ladiff_vae_decode.argtypes = [vp, vp, pi32, i32, i32, i32, vp, vp]
    if lib.ladiff_vae_encode is not None:
        lib.ladiff_vae_encode.argtypes = [vp, vp, pi32, i32, i32, vp, i32, vp, vp, vp, vp]
    lib.ladiff_feats2joints.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp, vp]
    lib.ladiff_linear_test.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp]
    lib.ladiff_linear_bench.argtypes = [vp, i32, i32, i32, i32, i32, i32, pf32, vp]
    lib.ladiff_ffn_test.argtypes = [vp, vp, i32, i32, vp, i32, i32, i32, vp, vp, pf32, vp]
    lib.ladiff_trace_read.argtypes = [vp, C.POINTER(C.c_uint64), i32, C.c_char_p, i32]
    lib.ladiff_trace_read.restype = C.c_int
    lib.ladiff_last_launch_count.argtypes = [vp]
    lib.ladiff_last_launch_count.restype = i64
    for fn in ("ladiff_create", "ladiff_set_weight", "ladiff_finalize_weights", "ladiff_diffusion_reverse",
               "ladiff_diffusion_reverse_ex", "ladiff_vae_encode", "ladiff_denoiser_forward", "ladiff_cfg_ddim_step", "ladiff_vae_decode", "ladiff_feats2joints",
               "ladiff_linear_test", "ladiff_linear_bench"):
        if getattr(lib, fn) is not None:
            getattr(lib, fn).restype = C.c_int
    _lib = lib
    return lib


EXPORTS = ("ladiff_abi_version", "ladiff_create", "ladiff_destroy", "ladiff_last_error", "ladiff_set_weight",
           "ladiff_finalize_weights", "ladiff_diffusion_reverse", "ladiff_diffusion_reverse_ex", "ladiff_denoiser_forward", "ladiff_cfg_ddim_step",
           "ladiff_vae_decode", "ladiff_vae_encode", "ladiff_feats2joints", "ladiff_linear_test", "ladiff_linear_bench", "ladiff_ffn_test", "ladiff_trace_read",
           "ladiff_last_launch_count")


def _i32(xs: Sequence[int]):
    return (C.c_int32 * len(xs))(*[int(x) for x in xs])


def _f32(xs: Sequence[float]):
    return (C.c_float * len(xs))(*[float(x) for x in xs])


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (ladiff_b200 has no CPU path)")
    return t.detach().to(torch.float32).contiguous()


class Engine:
    """One library handle = one (device, architecture config).  Shared by the denoiser / VAE mirror modules of a
    model so that their weights live in one place and the loop can run end to end on the device."""

    def __init__(self, nfeats: int = 263, max_it: int = 5, frame_per_latent: int = 48, max_frames: int = 196,
                 num_layers: int = 9, latent_dim: int = 256, num_heads: int = 4, ff_size: int = 1024,
                 text_dim: int = 768, use_cuda_graph: bool = True):
        self.lib = load_library()
        if not torch.cuda.is_available():
            raise RuntimeError("ladiff_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if os.environ.get("LADIFF_NO_GRAPH"):      # profiling / debugging: launch kernel by kernel
            use_cuda_graph = False
        self.cfg = LadiffConfig(nfeats, num_layers, latent_dim, num_heads, ff_size, text_dim, max_it, frame_per_latent,
                                max_frames, 1 if use_cuda_graph else 0)
        self.nfeats, self.max_it, self.max_frames = nfeats, max_it, max_frames
        self._h = C.c_void_p(0)
        st = self.lib.ladiff_create(C.byref(self.cfg), C.byref(self._h))
        if st != 0:
            msg = self.lib.ladiff_last_error(None).decode()
            raise (ValueError if st == -1 else RuntimeError)(f"ladiff_create: {msg}")
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._finalized = 0
        self._keep = []

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                self.lib.ladiff_destroy(self._h)
                self._h = C.c_void_p(0)
        except Exception:
            pass

    # -- errors -------------------------------------------------------------------------------
    def _check(self, st: int, what: str):
        if st == 0:
            return
        msg = self.lib.ladiff_last_error(self._h).decode()
        exc = {-1: ValueError, -2: KeyError, -3: RuntimeError, -4: RuntimeError}.get(st, RuntimeError)
        raise exc(f"{what}: {msg}")

    @property
    def last_launch_count(self) -> int:
        return int(self.lib.ladiff_last_launch_count(self._h))

    # -- weights --------------------------------------------------------------------------------
    def set_weights(self, state_dict: Dict[str, torch.Tensor], prefix: str):
        """state_dict keys WITHOUT the model-level prefix (as a sub-module's state_dict()); ``prefix`` is
        ``"denoiser."`` or ``"vae."`` (the reference's top-level names, modeltype/ladiff.py:90,109)."""
        for k, v in state_dict.items():
            t = _dev32(v.to(self.device), k)
            shape = (C.c_int64 * t.dim())(*t.shape)
            self._check(self.lib.ladiff_set_weight(self._h, (prefix + k).encode(), _ptr(t), shape, t.dim(), _stream()),
                        "set_weight")
        torch.cuda.current_stream().synchronize()
        self._finalized = 0

    def finalize(self, which: int):
        self._check(self.lib.ladiff_finalize_weights(self._h, which, _stream()), "finalize_weights")
        self._finalized |= which

    # -- compute ----------------------------------------------------------------------------------
    def diffusion_reverse(self, text_emb: torch.Tensor, lengths: Optional[Sequence[int]], noise: torch.Tensor,
                          timesteps: Sequence[int], c1: Sequence[float], c2: Sequence[float], guidance_scale: float,
                          mode: int, c3: Optional[Sequence[float]] = None, step_noise: Optional[torch.Tensor] = None,
                          seed: int = 0, rows: Optional[Sequence[int]] = None, autoregressive: bool = False) -> torch.Tensor:
        """The whole reverse loop (ladiff_diffusion_reverse[_ex]).  c3 / step_noise / seed: DDPM variance term (x' = c1 x +
        c2 eps + c3 noise; noise injected [n,B,T,256] or Philox(seed)); rows / autoregressive: the ARDIFF branch."""
        B = len(lengths) if lengths is not None else len(rows)
        text = _dev32(text_emb, "encoder_hidden_states").reshape(2 * B, -1)
        noise = _dev32(noise, "latents")
        if tuple(noise.shape) != (B, self.max_it, 256):
            raise ValueError(f"initial latents must be [{B},{self.max_it},256], got {tuple(noise.shape)}")
        z = torch.empty((self.max_it, B, 256), device=self.device, dtype=torch.float32)
        n = len(timesteps)
        if c3 is None and step_noise is None and rows is None and not autoregressive:
            self._check(self.lib.ladiff_diffusion_reverse(self._h, _ptr(text), _i32(lengths), B, _ptr(noise), n,
                                                          _i32(timesteps), _f32(c1), _f32(c2), float(guidance_scale),
                                                          mode, _ptr(z), _stream()), "diffusion_reverse")
            return z
        if step_noise is not None:
            step_noise = _dev32(step_noise, "step_noise")
            if tuple(step_noise.shape) != (n, B, self.max_it, 256):
                raise ValueError(f"step noise must be [{n},{B},{self.max_it},256], got {tuple(step_noise.shape)}")
        self._check(self.lib.ladiff_diffusion_reverse_ex(
            self._h, _ptr(text), _i32(lengths) if lengths is not None else None, _i32(rows) if rows is not None else None, B,
            _ptr(noise), n, _i32(timesteps), _f32(c1), _f32(c2), _f32(c3) if c3 is not None else None, _ptr(step_noise),
            C.c_uint64(int(seed) & (2 ** 64 - 1)), 1 if autoregressive else 0, float(guidance_scale), mode, _ptr(z), _stream()),
            "diffusion_reverse_ex")
        return z

    def denoiser_forward(self, sample: torch.Tensor, timestep: int, text_emb: torch.Tensor,
                         max_iter_elements: Sequence[int], mode: int) -> torch.Tensor:
        S = sample.shape[0]
        x = _dev32(sample, "sample")
        text = _dev32(text_emb, "encoder_hidden_states").reshape(S, -1)
        out = torch.empty_like(x)
        self._check(self.lib.ladiff_denoiser_forward(self._h, _ptr(x), int(timestep), _ptr(text),
                                                     _i32(max_iter_elements), S, mode, _ptr(out), _stream()),
                    "denoiser_forward")
        return out

    def cfg_ddim_step(self, noise_pred: torch.Tensor, latents: torch.Tensor, guidance_scale: float, c1: float,
                      c2: float) -> torch.Tensor:
        pred = _dev32(noise_pred, "noise_pred")
        lat = _dev32(latents, "latents").clone()
        self._check(self.lib.ladiff_cfg_ddim_step(self._h, _ptr(pred), _ptr(lat), lat.shape[0], float(guidance_scale),
                                                  float(c1), float(c2), _stream()), "cfg_ddim_step")
        return lat

    def vae_decode(self, z: torch.Tensor, lengths: Sequence[int], mode: int, max_len: Optional[int] = None) -> torch.Tensor:
        B = len(lengths)
        z = _dev32(z, "z")
        if tuple(z.shape) != (self.max_it, B, 256):
            raise ValueError(f"z must be [{self.max_it},{B},256], got {tuple(z.shape)}")
        max_len = int(max(lengths)) if max_len is None else int(max_len)
        out = torch.empty((B, max_len, self.nfeats), device=self.device, dtype=torch.float32)
        self._check(self.lib.ladiff_vae_decode(self._h, _ptr(z), _i32(lengths), B, max_len, mode, _ptr(out), _stream()),
                    "vae_decode")
        return out

    def vae_encode(self, feats: torch.Tensor, lengths: Sequence[int], mode: int, eps: Optional[torch.Tensor] = None):
        """(latent, mu, std), each [T, B, 256] (ladiff_vae_encode)."""
        f = _dev32(feats, "features")
        B, max_len, nf = f.shape
        if nf != self.nfeats or len(lengths) != B:
            raise ValueError(f"features must be [{len(lengths)}, max_len, {self.nfeats}], got {tuple(f.shape)}")
        eps = None if eps is None else _dev32(eps, "eps")
        if eps is not None and tuple(eps.shape) != (self.max_it, B, 256):
            raise ValueError(f"eps must be [{self.max_it},{B},256], got {tuple(eps.shape)}")
        lat, mu, std = (torch.empty((self.max_it, B, 256), device=self.device, dtype=torch.float32) for _ in range(3))
        self._check(self.lib.ladiff_vae_encode(self._h, _ptr(f), _i32(lengths), B, max_len, _ptr(eps), mode, _ptr(lat), _ptr(mu),
                                               _ptr(std), _stream()), "vae_encode")
        return lat, mu, std

    def feats2joints(self, feats: torch.Tensor, mean: torch.Tensor, std: torch.Tensor, njoints: int) -> torch.Tensor:
        f = _dev32(feats, "features")
        B, L, _ = f.shape
        out = torch.empty((B, L, njoints, 3), device=self.device, dtype=torch.float32)
        self._check(self.lib.ladiff_feats2joints(self._h, _ptr(f), _ptr(_dev32(mean, "mean")), _ptr(_dev32(std, "std")),
                                                 B, L, njoints, _ptr(out), _stream()), "feats2joints")
        return out

    def linear_test(self, A, W, bias=None, res=None, ln_g=None, ln_b=None, mod=None, epilogue="bias", mode=MODE_FP32):
        A, W = _dev32(A, "A"), _dev32(W, "W")
        M, K = A.shape
        N = W.shape[0]
        opt = [None if t is None else _dev32(t, "t") for t in (bias, res, ln_g, ln_b, mod)]
        out = torch.empty((M, N), device=self.device, dtype=torch.float32)
        self._check(self.lib.ladiff_linear_test(self._h, _ptr(A), _ptr(W), *[_ptr(t) for t in opt], M, N, K,
                                                EPI[epilogue], mode, _ptr(out), _stream()), "linear_test")
        return out

    def trace_read(self, max_launches: int = 16384, extra: bool = False):
        """[(name, start, wait_done, accum_ready, done)] in ns of the last diffusion_reverse (needs LADIFF_TRACE=1);
        extra=True appends the four kernel-specific stamps (slots 4..7; 0 where a kernel does not use them)."""
        buf = (C.c_uint64 * (8 * max_launches))()
        names = C.create_string_buffer(96 * max_launches)
        n = self.lib.ladiff_trace_read(self._h, buf, max_launches, names, 96)
        out = []
        for i in range(n):
            nm = names.raw[96 * i:96 * (i + 1)].split(b"\0", 1)[0].decode()
            rec = (nm, int(buf[8 * i]), int(buf[8 * i + 1]), int(buf[8 * i + 2]), int(buf[8 * i + 3]))
            if extra:
                rec = rec + tuple(int(buf[8 * i + k]) for k in range(4, 8))
            out.append(rec)
        return out

    def ffn_test(self, x: torch.Tensor, layer: int, mod: torch.Tensor, mode: int = MODE_BF16X3, fused=True, iters: int = 0):
        """(x3, s, ms) of the two feed-forward pairs of denoiser layer `layer` on rows x[M,256] (ladiff_ffn_test).
        fused: False = four separate fused linears, True = what the plans run (token-group kernel k_ffn_swap for
        M <= 1776, the separate linears above)."""
        x = x.contiguous().float()
        mod = mod.contiguous().float()
        x3, s = torch.empty_like(x), torch.empty_like(x)
        ms = C.c_float(0.0)
        self._check(self.lib.ladiff_ffn_test(self._h, _ptr(x), x.shape[0], layer, _ptr(mod), mode, int(fused), iters,
                                             _ptr(x3), _ptr(s), C.byref(ms), _stream()), "ffn_test")
        return x3, s, float(ms.value)

    def linear_bench(self, M: int, N: int, K: int, epilogue: str = "bias", mode: int = MODE_BF16X3, iters: int = 20) -> float:
        """average milliseconds per launch of one fused linear (device time, CUDA events)"""
        ms = C.c_float(0.0)
        self._check(self.lib.ladiff_linear_bench(self._h, M, N, K, EPI[epilogue], mode, iters, C.byref(ms), _stream()),
                    "linear_bench")
        return float(ms.value)
