"""Drop-in for the sampling surface of ``ladiff.models.modeltype.ladiff.LADIFF`` (reference lines 27-571):
``forward`` (:250-308), ``_diffusion_reverse`` LAD branch (:333-571) and ``gen_from_latent`` (:310-318).

Components are still created through ``instantiate_from_config(cfg.model.*)`` (:87-115) -- the reference's plugin
boundary -- so swapping the ``target:`` strings (``ladiff_b200.config.retarget``) is the whole integration.  The
N-step loop is ONE C-ABI call (``ladiff_diffusion_reverse``): CFG-doubled denoiser, CFG combine and the scheduler step
run on the device inside a captured CUDA graph; nothing returns to the host between steps.

Training, losses, metrics and the Lightning hooks are outside the hot path.
"""
from __future__ import annotations

import inspect
from typing import List, Optional

import torch
from torch import nn

from .config import instantiate_from_config
from .utils import max_iter_elements, remove_padding


class LADIFF(nn.Module):

    def __init__(self, cfg, datamodule, **kwargs):
        super().__init__()
        self.cfg = cfg
        self.stage = cfg.TRAIN.STAGE
        self.condition = cfg.model.condition
        self.is_vae = cfg.model.vae
        self.nfeats = cfg.DATASET.NFEATS
        self.njoints = cfg.DATASET.NJOINTS
        self.latent_dim = cfg.model.latent_dim
        self.guidance_scale = cfg.model.guidance_scale
        self.guidance_uncodp = cfg.model.guidance_uncondp
        self.datamodule = datamodule
        abl = cfg.TRAIN.ABLATION
        self.test_efficiency = abl.get("TEST_EFFICIENCY", False)
        self.max_it = abl.MAX_IT
        self.frame_per_latent = abl.FRAME_PER_LATENT
        self.joint_distro_fix = abl.get("JOINT_DISTRO_FIX", False)
        self.ARDIFF = cfg.get("ARDIFF", False)
        self.LAD = abl.get("LAD", True)
        if cfg.get("IDEA", "ard") != "ard" or not self.LAD or self.joint_distro_fix or self.test_efficiency:
            raise NotImplementedError("ladiff_b200 implements the LADiff sampling configurations: IDEA 'ard', ARDIFF False / True, "
                                      "LAD True, JOINT_DISTRO_FIX False, TEST_EFFICIENCY False "
                                      "(configs/config_ladiff_humanml3d.yaml:18-19,58-64)")
        self.motion_conditioning = cfg.model.get("motion_conditioning", "last")      # ARDIFF only (reference :52)
        nf = cfg.TRAIN.get("N_FRAMES", "None")
        self.nframes = None if nf in ("None", None) else nf                            # reference :58,63
        self.latent_dim_stage1 = cfg.model.get("latent_dim_stage1", None)
        if self.condition not in ("text", "text_uncond"):
            raise NotImplementedError("text-conditioned sampling only")
        try:
            self.vae_type = cfg.model.vae_type
        except (AttributeError, KeyError):
            self.vae_type = cfg.model.motion_vae.target.split(".")[-1].lower().replace("vae", "")

        self.text_encoder = instantiate_from_config(cfg.model.text_encoder)
        self.vae = instantiate_from_config(cfg.model.motion_vae)
        self.denoiser = instantiate_from_config(cfg.model.denoiser)
        self.scheduler = instantiate_from_config(cfg.model.scheduler)
        self.noise_scheduler = instantiate_from_config(cfg.model.noise_scheduler)
        for p in self.parameters():
            p.requires_grad = False
        self.do_classifier_free_guidance = self.guidance_scale > 1.0
        self.feats2joints = datamodule.feats2joints
        self.t2m_textencoder = self.t2m_moveencoder = self.t2m_motionencoder = None     # built on first t2m_eval
        self._shared_engine = None
        self._pipe_streams = None
        self.times: List[float] = []

    # -- engine plumbing ------------------------------------------------------------------------------------
    def _bind(self):
        """denoiser + VAE (+ feats2joints) share one library handle: one workspace, one stream of work."""
        if self._shared_engine is None:
            from ._lib import Engine
            self._shared_engine = Engine(nfeats=self.nfeats, max_it=self.max_it, frame_per_latent=self.frame_per_latent,
                                         max_frames=getattr(self.vae, "max_frames", 196))
            self.denoiser.bind_engine(self._shared_engine)
            self.vae.bind_engine(self._shared_engine)
            if hasattr(self.datamodule, "bind_engine"):
                self.datamodule.bind_engine(self._shared_engine)
        return self._shared_engine

    def set_precision(self, precision: str):
        """'bf16x3' (default, fp32-grade), 'bf16' (fastest) or 'fp32' (SIMT)."""
        from ._lib import MODES
        if precision not in MODES:
            raise ValueError(f"precision must be one of {sorted(MODES)}")
        self.denoiser.precision = precision
        self.vae.precision = precision

    # -- reference API ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, batch, latentwise_gen=None, plot_att_map=None):
        """batch {"text": List[str], "length": List[int]} -> List[Tensor[L_i, njoints, 3]]  (reference :250-308)."""
        if latentwise_gen is not None or plot_att_map is not None:
            raise NotImplementedError("latentwise_gen / plot_att_map are analysis options outside the sampling path")
        texts = batch["text"]
        lengths = batch["length"]
        if self.stage not in ("diffusion", "vae_diffusion"):
            raise NotImplementedError("stage 'vae' reconstructs from a given motion; not the sampling path")
        if self.do_classifier_free_guidance:
            uncond_tokens = [""] * len(texts)           # uncond FIRST (reference :258-264)
            if self.condition == "text":
                uncond_tokens.extend(texts)
            elif self.condition == "text_uncond":
                uncond_tokens.extend(uncond_tokens)
            texts = uncond_tokens
        text_emb = self.text_encoder(texts)
        z = self._diffusion_reverse(text_emb, lengths)
        feats_rst = self.vae.decode(z, lengths)
        return self._to_joints(feats_rst, lengths)

    def _to_joints(self, feats_rst, lengths):
        if getattr(self.datamodule, "accepts_cuda", False):
            joints = self.feats2joints(feats_rst.detach()).cpu()      # de-normalise + recover_from_ric on the GPU
        else:
            joints = self.feats2joints(feats_rst.detach().cpu())      # reference behaviour (:307)
        return remove_padding(joints, lengths)

    @torch.no_grad()
    def gen_from_latent(self, batch):
        """reference :310-318"""
        self._bind()
        feats_rst = self.vae.decode(batch["latent"], batch["length"])
        return self._to_joints(feats_rst, batch["length"])

    @torch.no_grad()
    def _diffusion_reverse(self, encoder_hidden_states, lengths=None, latents: Optional[torch.Tensor] = None,
                           generator: Optional[torch.Generator] = None, step_noise: Optional[torch.Tensor] = None):
        """encoder_hidden_states [2B,1,768] (uncond rows first), lengths List[int] -> latents [MAX_IT, B, 256] with rows
        >= ceil(L/48) exactly zero (reference :333-571, LAD branch).  ``latents=`` / ``generator=`` inject the initial noise
        the reference draws from the global RNG (:380-385)."""
        if lengths is None:
            raise ValueError("lengths are required on the length-aware path")
        engine = self._bind()
        bsz = encoder_hidden_states.shape[0]
        guidance = float(self.guidance_scale)
        if self.do_classifier_free_guidance:
            bsz = bsz // 2
        else:   # no guidance: eps = denoiser(cond); expressed as the CFG kernel with both halves equal and g = 1
            encoder_hidden_states = torch.cat([encoder_hidden_states] * 2)
            guidance = 1.0
        if len(lengths) != bsz:
            raise ValueError(f"{len(lengths)} lengths for {bsz} prompts")
        self.scheduler.set_timesteps(self.cfg.model.scheduler.num_inference_timesteps)   # :410
        eta = 0.0
        if "eta" in set(inspect.signature(self.scheduler.step).parameters.keys()):       # :415-417
            eta = self.cfg.model.scheduler.get("eta", 0.0)
        if not hasattr(self.scheduler, "fused_coefficients"):
            raise NotImplementedError("the fused loop needs a scheduler exposing fused_coefficients() (DDIM eta=0 / DDPM)")
        coef = self.scheduler.fused_coefficients(eta)
        ts, c1, c2 = coef[:3]
        c3 = coef[3] if len(coef) > 3 else None                 # DDPM: variance-noise coefficient per step
        self.denoiser.engine()       # weight sync
        lengths = [int(x) for x in lengths]
        extra = {}
        if c3 is not None:
            if step_noise is not None:
                extra = dict(c3=c3, step_noise=step_noise)
            else:   # the reference draws torch.randn inside scheduler.step; here: Philox stream seeded from torch's generator
                self._noise_calls = getattr(self, "_noise_calls", 0) + 1
                base = int(torch.randint(0, 2 ** 62, (1,)).item()) if generator is None else int(generator.initial_seed())
                extra = dict(c3=c3, seed=base + 0x9E3779B97F4A7C15 * self._noise_calls)
        if self.ARDIFF:
            return self._reverse_autoregressive(engine, encoder_hidden_states, lengths, bsz, latents, generator, ts, c1, c2,
                                                guidance, extra)
        if latents is None:
            latents = torch.randn((bsz, self.max_it, self.latent_dim[-1]), device=encoder_hidden_states.device,
                                  dtype=torch.float, generator=generator)
        latents = latents * self.scheduler.init_noise_sigma                     # :407 (masked rows are never read)
        return engine.diffusion_reverse(encoder_hidden_states, lengths, latents, ts, c1, c2, guidance, self.denoiser.mode, **extra)

    def _reverse_autoregressive(self, engine, text, lengths, bsz, latents, generator, ts, c1, c2, guidance, extra):
        """ARDIFF branch (reference :360-365, :419-467, :562-570): latent slot k is denoised with the N-step loop conditioned on
        the already-denoised slots (``enclat``: all of them for motion_conditioning 'full'/'middle', the last one for 'last');
        every iteration is one C-ABI call -- the conditioning latents ride along as fixed rows of the token sequence
        (ladiff_denoiser.py:247-248), no key-padding mask on this branch (max_iter_elements is not passed, :436-443)."""
        unit = self.latent_dim_stage1 if self.nframes is not None else self.frame_per_latent          # :343-356
        ar_iterations = -(-max(lengths) // int(unit))
        if ar_iterations > self.max_it:
            raise ValueError(f"{ar_iterations} autoregressive iterations exceed MAX_IT = {self.max_it}")
        D = self.latent_dim[-1]
        if latents is None:
            latents = torch.randn((bsz, ar_iterations, D), device=text.device, dtype=torch.float, generator=generator)   # :360-365
        if latents.shape[1] < ar_iterations:
            raise ValueError(f"initial latents need {ar_iterations} slots, got {latents.shape[1]}")
        latents = latents * self.scheduler.init_noise_sigma
        final = torch.zeros((bsz, self.max_it, D), device=text.device, dtype=torch.float)
        step_noise = extra.get("step_noise")
        for k in range(ar_iterations):
            if k == 0:
                ctx = final[:, :0]
            elif self.motion_conditioning in ("full", "middle"):
                ctx = final[:, :k]
            elif self.motion_conditioning == "last":
                ctx = final[:, k - 1:k]
            else:
                raise ValueError(f"motion_conditioning {self.motion_conditioning!r}")
            seq = torch.zeros((bsz, self.max_it, D), device=text.device, dtype=torch.float)
            seq[:, 0] = latents[:, k]
            seq[:, 1:1 + ctx.shape[1]] = ctx
            ex = dict(extra)
            if step_noise is not None:       # [ar_iterations, n, B, T, 256] when injected for the AR branch
                ex["step_noise"] = step_noise[k]
            elif "seed" in ex:
                ex["seed"] = ex["seed"] + k
            z = engine.diffusion_reverse(text, None, seq, ts, c1, c2, guidance, self.denoiser.mode,
                                         rows=[1 + ctx.shape[1]] * bsz, autoregressive=True, **ex)
            final[:, k] = z[0]
        out = final.permute(1, 0, 2).contiguous()                               # :466 -> [MAX_IT (zero-padded, :567-569), B, 256]
        for i, m in enumerate(max_iter_elements(lengths, self.frame_per_latent)):     # :562-566
            out[m:, i] = 0
        return out

    @torch.no_grad()
    def sample_features(self, encoder_hidden_states, lengths, latents=None, generator=None, max_len=None):
        """_diffusion_reverse + vae.decode without leaving the device: [B, max_len or max(lengths), nfeats] CUDA tensor."""
        z = self._diffusion_reverse(encoder_hidden_states, lengths, latents=latents, generator=generator)
        return self.vae.decode(z, lengths, max_len=max_len)

    # batch pairing of sample_stream: two consecutive batches of equal size run as ONE reverse-loop launch when together they make a
    # call the engine cuts into two independent chains (150 .. 354 prompts, each batch <= 177: ladiff_b200.cu pick_chains), so that
    # chain i is exactly the plan batch i would get alone
    PAIR_MIN_TOTAL, PAIR_MAX_TOTAL, PAIR_MAX_EACH = 150, 354, 177
    DECODE_MERGE_MIN_ROWS = 74 * 128      # B x max(lengths) above which the decoder's kernel selection no longer depends on the size

    def _pairable(self, a, b) -> bool:
        (ta, la, za), (tb, lb, zb) = a, b
        na, nb = len(la), len(lb)
        if na != nb:      # the engine cuts a call into equal chains: only equal batches map one batch to one chain (= its own plan)
            return False
        if not (self.PAIR_MIN_TOTAL <= na + nb <= self.PAIR_MAX_TOTAL and max(na, nb) <= self.PAIR_MAX_EACH):
            return False
        if (za is None) != (zb is None) or ta.device != tb.device or ta.dtype != tb.dtype or ta.shape[1:] != tb.shape[1:]:
            return False
        if (ta.shape[0] == 2 * na) != (tb.shape[0] == 2 * nb):       # both with or both without the CFG unconditional half
            return False
        return za is None or (za.shape[1:] == zb.shape[1:] and za.dtype == zb.dtype)

    @staticmethod
    def _merge_pair(a, b):
        (ta, la, za), (tb, lb, zb) = a, b
        na, nb = len(la), len(lb)
        if ta.shape[0] == 2 * na:                                      # [uncond | cond] halves (reference ladiff.py:259-265)
            text = torch.cat([ta[:na], tb[:nb], ta[na:], tb[nb:]], 0)
        else:
            text = torch.cat([ta, tb], 0)
        lat = None if za is None else torch.cat([za, zb], 0)
        return text, list(la) + list(lb), lat

    @torch.no_grad()
    def sample_stream(self, batches, pair: bool = True):
        """Pipelined ``sample_features`` over an iterable of ``(encoder_hidden_states, lengths, latents_or_None)``.

        Prompt batches are independent, and the two halves of the path load the GPU very differently: the 50-step reverse
        loop is a latency-bound chain of small launches (~110 of 148 SMs, mostly waiting), the LA-VAE decode is throughput
        work.  Two levels of overlap, both bit-identical to ``sample_features`` batch by batch (given the initial latents):

        * ``pair=True``: two consecutive batches that fit (see ``_pairable``; e.g. 2 x 128 prompts) are sampled by ONE
          ``_diffusion_reverse`` call, which the engine runs as two independent chains inside one CUDA graph -- each chain is
          exactly the launch sequence of its batch alone, and the chains fill each other's dependency gaps (B = 128 x 2:
          31.7 ms against 2 x 19.1 ms);
        * group i is decoded on a low-priority side stream while group i+1 already runs its reverse loop on a high-priority one.

        Yields the ``[B, max(lengths), nfeats]`` CUDA tensors in input order; a yielded tensor is ordered after its producer on
        the caller's current stream.  Per-batch latency is that of the group: callers that want one batch at a time use
        ``pair=False`` or ``sample_features``."""
        self._bind()
        cur = torch.cuda.current_stream()
        if getattr(self, "_pipe_streams", None) is None or self._pipe_streams[0].device != cur.device:
            self._pipe_streams = (torch.cuda.Stream(device=cur.device, priority=-1), torch.cuda.Stream(device=cur.device, priority=0))
        rev, dec = self._pipe_streams
        dec.wait_stream(cur)

        def groups():
            held = None
            for item in batches:
                if not pair:
                    yield [item]
                elif held is None:
                    held = item
                elif self._pairable(held, item):
                    yield [held, item]
                    held = None
                else:
                    yield [held]
                    held = item
            if held is not None:
                yield [held]

        pending = None
        for members in groups():
            text, lengths, lat = members[0] if len(members) == 1 else self._merge_pair(*members)
            rev.wait_stream(cur)                       # inputs were produced (copied / merged) on the caller's stream
            with torch.cuda.stream(rev):
                z = self._diffusion_reverse(text, lengths, latents=lat)
                ev = torch.cuda.Event()
                ev.record(rev)
            for t in (text, lat):
                if t is not None:
                    t.record_stream(rev)
            if pending is not None:                    # hand out group i only after group i+1 has been enqueued
                for feats, dev in pending:
                    cur.wait_event(dev)
                    feats.record_stream(cur)
                    yield feats
            dec.wait_event(ev)
            pending = []
            with torch.cuda.stream(dec):
                z.record_stream(dec)
                if len(members) > 1 and all(len(ln) * max(ln) > self.DECODE_MERGE_MIN_ROWS for _, ln, _ in members):
                    # both batches are past every size-dependent kernel choice of the decoder (> 74 row tiles of 128: whole-row
                    # LayerNorm GEMM, 256-wide tiles, tile FFN), and every decoder kernel is row- / sequence-local: ONE decode of
                    # the pair gives each batch bit for bit what its own decode gives, with better wave quantisation (392 row
                    # tiles on 148 SMs = 3 rounds instead of 2 x 2)
                    full = self.vae.decode(z, lengths)
                    dev = torch.cuda.Event()
                    dev.record(dec)
                    r0 = 0
                    for _, ln, _ in members:
                        part = full[r0:r0 + len(ln)]
                        if max(ln) != full.shape[1]:
                            part = part[:, :max(ln)].contiguous()
                            dev = torch.cuda.Event()
                            dev.record(dec)
                        r0 += len(ln)
                        pending.append((part, dev))
                else:
                    r0 = 0
                    for _, ln, _ in members:           # one decode per batch: exactly the launch sequence of sample_features
                        zi = z if len(members) == 1 else z[:, r0:r0 + len(ln)].contiguous()
                        r0 += len(ln)
                        feats = self.vae.decode(zi, ln)
                        dev = torch.cuda.Event()
                        dev.record(dec)
                        pending.append((feats, dev))
        if pending is not None:
            for feats, dev in pending:
                cur.wait_event(dev)
                feats.record_stream(cur)
                yield feats

    @torch.no_grad()
    def t2m_eval(self, batch, is_mm: bool = False, latents: Optional[torch.Tensor] = None, eps: Optional[torch.Tensor] = None):
        """Evaluation glue (reference :1111-1282): stage 'diffusion' samples from the texts, stage 'vae' reconstructs the ground
        truth through ``vae.encode`` -> ``vae.decode``; both then go through feats2joints, the T2M renormalisation and the T2M
        co-embedding evaluators.  Returns the reference's dict (m_ref, m_rst, lat_t, lat_m, lat_rm, joints_ref, joints_rst).
        ``latents`` / ``eps`` inject the random draws (initial noise / rsample) for parity tests."""
        texts = list(batch["text"])
        motions = batch["motion"].detach().clone()
        lengths = [int(x) for x in batch["length"]]
        word_embs = batch["word_embs"].detach().clone()
        pos_ohot = batch["pos_ohot"].detach().clone()
        text_lengths = batch["text_len"].detach().clone()
        if is_mm:                                                                          # :1123-1134
            r = self.cfg.TEST.MM_NUM_REPEATS
            texts, lengths = texts * r, lengths * r
            motions, word_embs, pos_ohot, text_lengths = (t.repeat_interleave(r, dim=0) for t in (motions, word_embs, pos_ohot, text_lengths))
        self._bind()
        if self.stage in ("diffusion", "vae_diffusion"):                                   # :1136-1147
            if self.do_classifier_free_guidance:
                uncond = [""] * len(texts)
                uncond.extend(texts if self.condition == "text" else uncond)
                text_in = uncond
            else:
                text_in = texts
            z = self._diffusion_reverse(self.text_encoder(text_in), lengths, latents=latents)
        elif self.stage == "vae":                                                          # :1152-1156
            z, _, _ = self.vae.encode(motions, lengths, eps=eps)
            if self.condition == "text_uncond":
                z = torch.randn_like(z)
        else:
            raise NotImplementedError(f"stage {self.stage!r}")
        feats_rst = self.vae.decode(z, lengths)                                            # :1203
        # match the ground-truth length (:1218-1229; decode already pads to max(lengths) with zeros)
        max_len = max(lengths)
        if feats_rst.shape[1] != max_len:
            out = feats_rst.new_zeros((feats_rst.shape[0], max_len, feats_rst.shape[2]))
            n = min(max_len, feats_rst.shape[1])
            out[:, :n] = feats_rst[:, :n]
            feats_rst = out
        motions = motions.to(feats_rst.device)
        on_gpu = getattr(self.datamodule, "accepts_cuda", False)
        joints_rst = self.feats2joints(feats_rst if on_gpu else feats_rst.cpu())         # :1245-1246
        joints_ref = self.feats2joints(motions if on_gpu else motions.cpu())
        feats_rst = self.datamodule.renorm4t2m(feats_rst)                                  # :1250-1251
        motions = self.datamodule.renorm4t2m(motions)
        m_lens = torch.tensor(lengths, device=motions.device)                              # :1254-1261
        align_idx = torch.argsort(m_lens, descending=True, stable=True)
        motions, feats_rst, m_lens = motions[align_idx], feats_rst[align_idx], m_lens[align_idx]
        m_lens = torch.div(m_lens, self.cfg.DATASET.HUMANML3D.UNIT_LEN, rounding_mode="floor")
        if getattr(self, "t2m_moveencoder", None) is None:
            from .evaluators import build_t2m_evaluators
            self.t2m_textencoder, self.t2m_moveencoder, self.t2m_motionencoder = (
                m.to(motions.device) for m in build_t2m_evaluators(self.cfg, self.nfeats))
        recons_emb = self.t2m_motionencoder(self.t2m_moveencoder(feats_rst[..., :-4]), m_lens)      # :1263-1266
        motion_emb = self.t2m_motionencoder(self.t2m_moveencoder(motions[..., :-4]), m_lens)
        text_emb = self.t2m_textencoder(word_embs.to(motions.device), pos_ohot.to(motions.device), text_lengths)[align_idx]   # :1269-1270
        return {"m_ref": motions, "m_rst": feats_rst, "lat_t": text_emb, "lat_m": motion_emb, "lat_rm": recons_emb,
                "joints_ref": joints_ref, "joints_rst": joints_rst}
