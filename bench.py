#!/usr/bin/env python
"""Benchmark of the LADiff sampling hot path (BASELINE.json metric: motion sequences / s, 50-step DDIM + CFG 7.5,
196 frames, followed by the LA-VAE decode).

    python bench.py --gpus 1 --steps 10 --warmup 3                 # own arm (CUDA path through the public API)
    python bench.py --impl reference --steps 2 --warmup 1          # reference arm: the CPU oracle on the host cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W                       # one rank per GPU, weak scaling

A "step" = one batch of 128 synthetic prompts (random-init weights, injected noise) through
``LADIFF._diffusion_reverse`` + ``vae.decode``.  CLIP is timed separately (``clip_ms``), as north_star asks.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU = 128
FRAMES = 196
STEPS_DDIM = 50
GUIDANCE = 7.5
NFEATS = 263


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures (profiles/)
NCU_DRAM_BYTES = {"k_ffn_swap<2>": 7174912, "k_ffn_cluster<2>": 7165952}


_REAL_STDOUT = None


def quiet_stdout():
    """stdout carries exactly ONE JSON line: anything a library prints there (NCCL's version banner, ...) goes to stderr."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        print(line, flush=True)
    else:
        os.write(_REAL_STDOUT, (line + "\n").encode())


def lin(i, o):
    return 2.0 * i * o


def algorithmic_flops(lengths, n_steps=STEPS_DDIM, nfeats=NFEATS, fpl=48, T=5):
    """SURVEY.md 8d contract figure: useful work only (valid rows / frames; hoisted or dead computations excluded)."""
    den = dec = 0.0
    for L in lengths:
        m = min(T, -(-L // fpl))
        f_row = 9 * (lin(256, 768) + lin(256, 256) + lin(256, 1024) + lin(1024, 256) + 2 * ((m + 2) * 256 * 2)
                     + lin(256, 256) + lin(256, 1024) + lin(1024, 256) + lin(256, 256)) + 4 * lin(512, 256)
        den += 2 * m * f_row * n_steps
        d_row = 9 * (lin(256, 768) + lin(256, 256) + 2 * (L * 256 * 2) + lin(256, 256) + 2 * (m * 256 * 2)
                     + lin(256, 256) + lin(256, 1024) + lin(1024, 256)) + 4 * lin(512, 256)
        dec += L * (d_row + lin(256, nfeats)) + 9 * m * lin(256, 512)
    return den, dec


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"], "hbm_gbs": d["hbm_gbs"],
                "source": "MEASURED_PEAKS.json"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
def cpu_reference_step(O, sd, text, noise, lengths, den_steps):
    """One bounded sample of the reference's CPU path (oracle restatement): `den_steps` of the 50 denoiser steps, scaled
    linearly (every step is identical work), plus one full decode.  Returns seconds for the full 50-step workload."""
    import torch
    t0 = time.perf_counter()
    mie = O.max_iter_elements_of(lengths)
    lat = O.initial_latents(noise, lengths)
    acp = O.ddim_alphas_cumprod()
    ts = O.ddim_timesteps(STEPS_DDIM)
    mie2 = torch.cat([mie] * 2)
    for t in ts[:den_steps]:
        pred = O.denoiser_forward(sd, torch.cat([lat] * 2), torch.tensor(int(t)), text, mie2)
        u, c = pred.chunk(2)
        lat = O.ddim_step(u + GUIDANCE * (c - u), int(t), lat, acp, STEPS_DDIM)
    t_den = time.perf_counter() - t0
    t0 = time.perf_counter()
    O.vae_decode(sd, O.initial_latents(noise, lengths).permute(1, 0, 2).contiguous(), lengths)
    t_dec = time.perf_counter() - t0
    return t_den * (STEPS_DDIM / den_steps) + t_dec, t_den, t_dec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import ladiff_oracle as O
    torch.set_grad_enabled(False)
    try:   # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm runs on rank 0 alone and gets all host threads
        torch.set_num_threads(len(os.sched_getaffinity(0)))
    except Exception:
        pass
    cores = torch.get_num_threads()
    sd = O.make_state_dict(1234, NFEATS, perturb=False)
    text, noise, lengths = O.synthetic_inputs(B_PER_GPU, seed=1234, ragged=False, fixed_len=FRAMES)
    den_steps = 5
    for _ in range(args.warmup):
        cpu_reference_step(O, sd, text, noise, lengths, 1)
    times = [cpu_reference_step(O, sd, text, noise, lengths, den_steps)[0] for _ in range(args.steps)]
    sec = sum(times) / len(times)
    val = B_PER_GPU / sec
    sample = f"B={B_PER_GPU} L={FRAMES}: {den_steps} of {STEPS_DDIM} CFG denoiser steps timed and scaled x{STEPS_DDIM // den_steps} + one full decode, fp32 torch CPU"
    emit(json.dumps({
        "impl": "reference", "metric": "motion sequences/sec (50-step DDIM+CFG, 196 frames)", "value": val, "unit": "seq/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"LA-DDPM sampling batch {B_PER_GPU}, {STEPS_DDIM}-step DDIM + CFG {GUIDANCE}, {FRAMES} frames, + LA-VAE decode; CPU"},
        "cpu_baseline": {"value": val, "unit": "seq/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "seq/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------------------------------
def run_own(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # keep stdout to the ONE JSON line: NCCL prints its version banner there at NCCL_DEBUG=VERSION
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.set_grad_enabled(False)

    import ladiff_b200 as L
    from ladiff_b200.data import SyntheticDataModule
    from ladiff_b200.modeltype import LADIFF

    cfg = L.default_config("humanml3d", num_inference_timesteps=STEPS_DDIM)
    torch.manual_seed(1234)                                  # configs/base.yaml:2 SEED_VALUE
    model = LADIFF(cfg, SyntheticDataModule(NFEATS, 22)).to(dev).eval()   # random-init weights of the reference architecture
    model.set_precision(args.mode)
    B = args.batch
    g = torch.Generator().manual_seed(1234 + rank)
    lengths = [FRAMES] * B
    text_h = torch.randn((2 * B, 1, 768), generator=g).pin_memory()
    noise_h = torch.randn((B, 5, 256), generator=g).pin_memory()
    text_d, noise_d = text_h.to(dev), noise_h.to(dev)
    out_h = torch.empty((B, FRAMES, NFEATS), dtype=torch.float32).pin_memory()
    gather = [torch.empty((B, FRAMES, NFEATS), device=dev) for _ in range(world)] if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)   # > 126 MB L2
    eng = model._bind()

    def step_resident():
        z = model._diffusion_reverse(text_d, lengths, latents=noise_d)
        n1 = eng.last_launch_count
        feats = model.vae.decode(z, lengths)
        n2 = eng.last_launch_count
        if world > 1:
            dist.all_gather(gather, feats)                    # the only collective: motions over NVLink
        return feats, n1 + n2

    def step_e2e():
        t = text_h.to(dev, non_blocking=True)
        nz = noise_h.to(dev, non_blocking=True)
        z = model._diffusion_reverse(t, lengths, latents=nz)
        feats = model.vae.decode(z, lengths)
        if world > 1:
            dist.all_gather(gather, feats)
        out_h.copy_(feats, non_blocking=True)
        return feats

    def timed(fn, K, W):
        for _ in range(W):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        evs = []
        for _ in range(K):
            flush.fill_(1.0)                                  # L2 flush between timed iterations (outside the events)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- pipelined over the K steps (LADIFF.sample_stream): the K batches are independent, batch i is decoded on a
    # low-priority stream while batch i+1 runs its reverse loop.  Same work, same results; one event pair around the K steps.
    def batches(n, host):
        for _ in range(n):
            flush.fill_(1.0)                                  # L2 flush between iterations (inside the timed region here)
            if host:
                yield text_h.to(dev, non_blocking=True), lengths, noise_h.to(dev, non_blocking=True)
            else:
                yield text_d, lengths, noise_d

    def run_pipe(n, host):
        for feats in model.sample_stream(batches(n, host)):
            if world > 1:
                dist.all_gather(gather, feats)                # the only collective: motions over NVLink
            if host:
                out_h.copy_(feats, non_blocking=True)

    def timed_pipe(K, W, host):
        run_pipe(W, host)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_pipe(K, host)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    W = max(args.warmup, 3)
    _, launches = step_resident()                             # builds plans / graphs
    pipelined = not args.no_pipeline
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_seq = timed(lambda: step_resident(), args.steps, W)    # one batch at a time (latency view)
    ms_total = timed_pipe(args.steps, W, False) if pipelined else ms_seq
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed_pipe(args.steps, W, True) if pipelined else timed(step_e2e, args.steps, W)
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1e3)
    e2e = world * B * args.steps / (ms_e2e / 1e3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- extras (rank 0, outside the headline timing) ---------------------------------------------------------------------------
    peaks = measured_peaks()
    den_f, dec_f = algorithmic_flops(lengths)
    flops = den_f + dec_f
    achieved = flops / (ms_step / 1e3) / 1e12
    extra = {}
    if world == 1 and not args.quick:
        # component split
        def t_ms(fn, K=5):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(K):
                fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / K
        z = model._diffusion_reverse(text_d, lengths, latents=noise_d)
        extra["reverse_ms"] = t_ms(lambda: model._diffusion_reverse(text_d, lengths, latents=noise_d))
        extra["decode_ms"] = t_ms(lambda: model.vae.decode(z, lengths))
        # dominant kernel (the fused tcgen05 linear) timed alone, CUDA events inside the library
        from ladiff_b200._lib import MODES
        mode = MODES[args.mode]
        kern = {}
        for name, (M, N, K_, epi) in {"dec_ffn1_gelu": (B * FRAMES, 1024, 256, "gelu"), "dec_ffn2_ln": (B * FRAMES, 256, 1024, "ln"),
                                      "den_ffn1_gelu": (2 * B * 5, 1024, 256, "gelu"), "den_ffn2_ln": (2 * B * 5, 256, 1024, "ln"),
                                      "den_qkv": (2 * B * 5, 768, 256, "bias")}.items():
            ms = eng.linear_bench(M, N, K_, epi, mode, 20)
            tf = 2.0 * M * N * K_ / (ms / 1e3) / 1e12
            kern[name] = {"M": M, "N": N, "K": K_, "us": ms * 1e3, "tflops": tf, "frac_of_bf16_burst": tf / peaks["bf16_burst"]}
        extra["kernels"] = kern
        # the dominant kernel of the step (43 % of the launch-list time, profiles/r01i_launches_bf16x3.summary.txt): the fused
        # feed-forward pairs of one denoiser layer (k_ffn_swap), timed alone with CUDA events inside the library on its stream
        xk = torch.randn((2 * B * 5, 256), generator=g).to(dev)
        mk = (0.3 * torch.randn((512,), generator=g)).to(dev)
        _, _, ms_k = eng.ffn_test(xk, 4, mk, mode=mode, fused=True, iters=200)
        fl_k = 2 * B * 5 * 2 * (lin(256, 1024) + lin(1024, 256))
        extra["_dominant"] = {"kernel": "k_ffn_swap<2>" if args.mode == "bf16x3" else "k_ffn_swap<1>", "us_per_launch": ms_k * 1e3,
                              "flops_per_launch": fl_k, "launches_per_step": 9 * STEPS_DDIM}
        # other precision modes, same workload
        for m in ("bf16", "bf16x3", "fp32"):
            if m == args.mode or (m == "fp32" and args.steps < 3):
                continue
            model.set_precision(m)
            step_resident()
            k = 3 if m != "fp32" else 1
            extra[f"value_{m}"] = B * k / (timed(lambda: step_resident(), k, 1) / 1e3)
        model.set_precision(args.mode)
        # large-batch throughput (the 8192-prompt sweep's per-GPU micro-batch): fills all 148 SMs
        Bb = 1024
        tb = torch.randn((2 * Bb, 1, 768), generator=g).to(dev)
        nb = torch.randn((Bb, 5, 256), generator=g).to(dev)
        lb = [FRAMES] * Bb
        fb = lambda: model.vae.decode(model._diffusion_reverse(tb, lb, latents=nb), lb)
        ms_b = t_ms(fb, 2)
        fl_b = sum(algorithmic_flops(lb))
        extra["batch1024"] = {"value": Bb / (ms_b / 1e3), "unit": "seq/s", "ms": ms_b,
                              "roofline_frac": fl_b / (ms_b / 1e3) / 1e12 / peaks["bf16_sustained"]}
        # BASELINE configs 2 / 3, ragged variants (lengths = 4 * U[10,50), seed 1234): decode only and full sampling
        import numpy as np
        lr = [int(x) for x in (np.random.default_rng(1234).integers(10, 50, size=B) * 4)]
        zr = model._diffusion_reverse(text_d, lr, latents=noise_d)
        den_r, dec_r = algorithmic_flops(lr)
        ms_dr = t_ms(lambda: model.vae.decode(zr, lr))
        ms_sr = t_ms(lambda: model.vae.decode(model._diffusion_reverse(text_d, lr, latents=noise_d), lr))
        extra["ragged"] = {"lengths": "4*U[10,50) seed 1234 (mean %.1f frames)" % (sum(lr) / len(lr)),
                           "decode_only_ms": ms_dr, "decode_only_seq_s": B / (ms_dr / 1e3),
                           "decode_only_tflops": dec_r / (ms_dr / 1e3) / 1e12,
                           "sampling_ms": ms_sr, "sampling_seq_s": B / (ms_sr / 1e3),
                           "sampling_tflops": (den_r + dec_r) / (ms_sr / 1e3) / 1e12}
        # CLIP, timed separately
        texts = [""] * B + [f"a person walks forward then turns {i}" for i in range(B)]
        model.text_encoder(texts); torch.cuda.synchronize()
        t0 = time.perf_counter(); model.text_encoder(texts); torch.cuda.synchronize()
        extra["clip_ms"] = (time.perf_counter() - t0) * 1e3
        # CPU baseline: the reference path (oracle port) on this box's host cores, bounded sample
        from oracle import ladiff_oracle as O
        try:
            torch.set_num_threads(len(os.sched_getaffinity(0)))     # all host threads, whatever OMP_NUM_THREADS the launcher exported
        except Exception:
            pass
        sd = O.make_state_dict(1234, NFEATS, perturb=False)
        ctext, cnoise, clen = O.synthetic_inputs(B, seed=1234, ragged=False, fixed_len=FRAMES)
        cpu_reference_step(O, sd, ctext, cnoise, clen, 1)
        sec, _, _ = cpu_reference_step(O, sd, ctext, cnoise, clen, 5)
        cpu = {"value": B / sec, "unit": "seq/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"B={B} L={FRAMES}: 5 of 50 CFG denoiser steps timed and scaled x10 + one full decode, fp32 torch CPU (oracle/ladiff_oracle.py)"}
    else:
        cpu = None

    out = {
        "metric": "motion sequences/sec (50-step DDIM+CFG, 196 frames)", "value": value, "unit": "seq/s", "n_gpus": world,
        "steps": args.steps, "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None,
        "dtype": {"bf16x3": "bf16 hi/lo split x3 products, fp32 accumulate (fp32-grade, parity 1e-3)", "bf16": "bf16, fp32 accumulate",
                  "fp32": "f32"}[args.mode],
        "data": "synthetic",
        "config": {"workload": f"LA-DDPM sampling batch {B} per GPU, {STEPS_DDIM}-step DDIM + CFG {GUIDANCE}, {FRAMES} frames, + LA-VAE decode to {NFEATS}-d features",
                   "weights": "random-init (reference initialiser families), seed 1234", "mode": args.mode,
                   "l2": "L2 flushed (256 MiB write) between timed iterations", "collective": "all_gather of motions per step" if world > 1 else "none",
                   "schedule": ("pipelined over the K steps (LADIFF.sample_stream: decode of batch i on a low-priority stream under the reverse "
                                "loop of batch i+1)") if pipelined else "one batch at a time"},
        "latency_ms_per_batch": ms_seq / args.steps, "value_sequential": world * B * args.steps / (ms_seq / 1e3),
        "e2e": {"value": e2e, "unit": "seq/s", "h2d_bytes_per_step": int(text_h.numel() * 4 + noise_h.numel() * 4),
                "d2h_bytes_per_step": int(out_h.numel() * 4), "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches * args.steps),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["bf16_sustained"], "traffic": None, "scope": "whole step",
                     "note": f"algorithmic FLOPs per step {flops / 1e12:.3f} T (denoiser {den_f / 1e12:.3f} + decoder {dec_f / 1e12:.3f}; SURVEY.md 8d) / step time; peak = bf16 sustained of {peaks['source']}"},
        "cpu_baseline": cpu,
    }
    dom = extra.pop("_dominant", None)
    if dom:
        # roofline of the DOMINANT KERNEL (timed alone -> burst peak); the whole-step figure moves to roofline.path
        ach_k = dom["flops_per_launch"] / (dom["us_per_launch"] * 1e-6) / 1e12
        path = dict(out["roofline"])
        out["roofline"] = {
            "bound": "tensor", "achieved": ach_k, "peak": peaks["bf16_burst"], "unit": "TFLOP/s", "frac": ach_k / peaks["bf16_burst"],
            "traffic": NCU_DRAM_BYTES.get(dom["kernel"]), "kernel": dom["kernel"], "us_per_launch": dom["us_per_launch"],
            "launches_per_step": dom["launches_per_step"],
            "share_of_step": dom["us_per_launch"] * 1e-3 * dom["launches_per_step"] / (ms_seq / args.steps),
            "note": ("algorithmic FLOPs per launch = 1280 rows x 4 x lin(256,1024) (the two feed-forward pairs of one denoiser layer, "
                     "SURVEY.md 8d rows 'sa ReLU-FFN' + 'GELU FFN'); bf16x3 issues 3x these on the tensor pipe; duration = CUDA events "
                     f"around 200 back-to-back launches; peak = bf16 burst of {peaks['source']}; traffic = dram read+write per launch "
                     "of the ncu --set full capture in profiles/r01h_ncu_full_layer_kernels.txt"),
            "path": path}
    out.update(extra)
    emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--mode", default="bf16x3", choices=["bf16x3", "bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU)
    ap.add_argument("--no-pipeline", action="store_true", help="time one batch at a time instead of LADIFF.sample_stream")
    ap.add_argument("--quick", action="store_true", help="skip the extra measurements (kernel table, other modes, CPU baseline)")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
