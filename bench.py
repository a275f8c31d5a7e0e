#!/usr/bin/env python
"""Benchmark of the LADiff sampling hot path (BASELINE.json metric: motion sequences / s, 50-step DDIM + CFG 7.5,
196 frames, followed by the LA-VAE decode).

    python bench.py --gpus 1 --steps 10 --warmup 3                 # own arm (CUDA path through the public API)
    python bench.py --impl reference --steps 2 --warmup 1          # reference arm: the CPU oracle on the host cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W                       # one rank per GPU, weak scaling

    python bench.py --sweep 8192 [--micro 1024]                     # BASELINE config 5: N prompts sharded over the ranks (strong scaling)

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


METRIC = "motion sequences/sec (50-step DDIM+CFG, 196 frames)"


def workload_config(batch=B_PER_GPU):
    """The SAME config object in both arms (the driver compares them): what is computed, not how."""
    return {"workload": f"LA-DDPM sampling batch {batch} per GPU, {STEPS_DDIM}-step DDIM + CFG {GUIDANCE}, {FRAMES} frames, + LA-VAE decode to {NFEATS}-d features",
            "weights": "random-init (reference initialiser families), seed 1234", "precision": "fp32-grade (1e-3 max-abs parity contract on decoded features)"}


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full captures (profiles/)
NCU_DRAM_BYTES = {"k_ffn_swap<2>": 7174912}


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
def cpu_reference_run(O, sd, text, noise, lengths, den_steps=STEPS_DDIM):
    """The reference's CPU path (oracle restatement, un-hoisted) on one batch: ALL `den_steps` CFG denoiser steps + DDIM + the
    decode.  Returns (seconds, decoded features)."""
    t0 = time.perf_counter()
    z = O.diffusion_reverse(sd, text, lengths, noise, den_steps, GUIDANCE)
    feats = O.vae_decode(sd, z, lengths)
    return time.perf_counter() - t0, feats


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
    for _ in range(max(1, args.warmup)):
        cpu_reference_run(O, sd, text, noise, lengths, 2)                    # warm-up steps (threads, allocator): 2 denoiser steps each
    times = [cpu_reference_run(O, sd, text, noise, lengths)[0] for _ in range(args.steps)]
    times.sort()
    sec = sum(times) / len(times)
    val = B_PER_GPU / sec
    sample = (f"B={B_PER_GPU} L={FRAMES}: every step runs all {STEPS_DDIM} CFG denoiser steps + DDIM + one full decode "
              f"(nothing extrapolated), fp32 torch CPU, {cores} threads (oracle/ladiff_oracle.py)")
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "seq/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(),
        "p50_latency_ms": times[len(times) // 2] * 1e3,
        "cpu_baseline": {"value": val, "unit": "seq/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "seq/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# --------------------------------------------------------------------------------------------------
def setup(args):
    """One process per GPU: device, NCCL group (world > 1) and the model with random-init weights of the reference architecture."""
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
    return torch, dist, rank, world, local, dev, model


def run_sweep(args):
    """BASELINE config 5: N synthetic prompts sharded over the ranks (``parallel.shard_range``: contiguous, balanced), each rank
    samples its shard in micro-batches through ``LADIFF.sample_stream`` (host inputs, pinned), and ONE all-gather
    (``parallel.gather_motions``) returns every motion to every rank.  Strong scaling: total work fixed as the rank count grows.
    Timed from the first H2D copy to the end of the all-gather, device events, max over ranks."""
    torch, dist, rank, world, local, dev, model = setup(args)
    from ladiff_b200.parallel import gather_motions, shard_range
    N, micro = args.sweep, args.micro
    s0, e0 = shard_range(N, rank, world)
    n_local = e0 - s0
    g = torch.Generator().manual_seed(1234 + rank)
    nb = -(-n_local // micro)
    sizes = [min(micro, n_local - i * micro) for i in range(nb)]
    # one pinned micro-batch of prompts per distinct size, re-used (the content does not change the work)
    host = {b: (torch.randn((2 * b, 1, 768), generator=g).pin_memory(), torch.randn((b, 5, 256), generator=g).pin_memory()) for b in set(sizes)}
    local_out = torch.zeros((n_local, FRAMES, NFEATS), device=dev)

    def batches():
        for b in sizes:
            t, nz = host[b]
            yield t.to(dev, non_blocking=True), [FRAMES] * b, nz.to(dev, non_blocking=True)

    def run():
        i = 0
        for feats in model.sample_stream(batches()):
            local_out[i:i + feats.shape[0]].copy_(feats)
            i += feats.shape[0]
        return gather_motions(local_out, [FRAMES] * n_local, N, max_len=FRAMES)

    for b in set(sizes):                                      # warm-up: plans / graphs of every micro-batch size, NCCL communicator
        t, nz = host[b]
        for _ in range(2):
            model.vae.decode(model._diffusion_reverse(t.to(dev), [FRAMES] * b, latents=nz.to(dev)), [FRAMES] * b)
    if world > 1:
        dist.all_reduce(torch.zeros(1, device=dev))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    motions, lens = run()
    eb.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ea.elapsed_time(eb)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ok = tuple(motions.shape) == (N, FRAMES, NFEATS) and len(lens) == N and bool(torch.isfinite(motions[::97]).all())
    if rank == 0:
        peaks = measured_peaks()
        fl = sum(algorithmic_flops([FRAMES] * N))
        emit(json.dumps({
            "metric": METRIC, "value": N / (ms / 1e3), "unit": "seq/s", "n_gpus": world, "steps": nb, "warmup": 2,
            "ms_per_step": ms / nb, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.mode, "data": "synthetic",
            "config": {"workload": f"throughput sweep: {N} synthetic prompts sharded over {world} GPU(s), micro-batch {micro}, "
                                   f"{STEPS_DDIM}-step DDIM + CFG {GUIDANCE}, {FRAMES} frames, decode to {NFEATS}-d, one all-gather of all motions"},
            "sweep": {"prompts": N, "micro_batch": micro, "prompts_per_rank": n_local, "micro_batches_per_rank": nb, "total_ms": ms,
                      "all_motions_on_every_rank": ok, "gathered_bytes": int(motions.numel() * 4)},
            "e2e": {"value": N / (ms / 1e3), "unit": "seq/s", "h2d_bytes_per_step": int(micro * (2 * 768 + 5 * 256) * 4), "d2h_bytes_per_step": 0},
            "roofline": {"bound": "tensor", "achieved": fl / (ms / 1e3) / 1e12, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                         "frac": fl / (ms / 1e3) / 1e12 / peaks["bf16_sustained"], "traffic": None, "scope": "whole sweep"},
            "clocks": clocks}))
    if world > 1:
        dist.destroy_process_group()


def run_own(args):
    torch, dist, rank, world, local, dev, model = setup(args)
    import ladiff_b200 as L
    from ladiff_b200.data import SyntheticDataModule
    from ladiff_b200.modeltype import LADIFF
    B = args.batch
    g = torch.Generator().manual_seed(1234 + rank)
    lengths = [FRAMES] * B
    text_h = torch.randn((2 * B, 1, 768), generator=g).pin_memory()
    noise_h = torch.randn((B, 5, 256), generator=g).pin_memory()
    text_d, noise_d = text_h.to(dev), noise_h.to(dev)
    out_h = torch.empty((B, FRAMES, NFEATS), dtype=torch.float32).pin_memory()
    gather = torch.empty((world * B, FRAMES, NFEATS), device=dev) if world > 1 else None   # every rank's [B, 196, 263] slot
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)   # > 126 MB L2
    eng = model._bind()

    def step_resident():
        z = model._diffusion_reverse(text_d, lengths, latents=noise_d)
        n1 = eng.last_launch_count
        feats = model.vae.decode(z, lengths)
        n2 = eng.last_launch_count
        if world > 1:
            dist.all_gather_into_tensor(gather, feats)        # the only collective: motions over NVLink
        return feats, n1 + n2

    def step_e2e():
        t = text_h.to(dev, non_blocking=True)
        nz = noise_h.to(dev, non_blocking=True)
        z = model._diffusion_reverse(t, lengths, latents=nz)
        feats = model.vae.decode(z, lengths)
        if world > 1:
            dist.all_gather_into_tensor(gather, feats)
        out_h.copy_(feats, non_blocking=True)
        return feats

    step_ms = []           # per-batch device times of the last timed() call (sorted)

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
        per = sorted(a.elapsed_time(b) for a, b in evs)
        step_ms[:] = per
        t = torch.tensor([sum(per)], device=dev, dtype=torch.float64)
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

    d2h = torch.cuda.Stream(device=dev)      # result read-back on its own stream: the next batch's H2D copies (caller's stream,
                                             # which sample_stream orders the reverse loop after) must not queue behind it

    comm = torch.cuda.Stream(device=dev) if world > 1 else None   # same reason for the all-gather: off the caller's stream

    def run_pipe(n, host):
        cur = torch.cuda.current_stream()
        for feats in model.sample_stream(batches(n, host)):
            if world > 1:
                # the only collective: motions over NVLink, straight into the preallocated [world * B, 196, 263] buffer, on its
                # own stream, while the NEXT pair's reverse loop already runs on the high-priority stream
                comm.wait_stream(cur)
                with torch.cuda.stream(comm):
                    dist.all_gather_into_tensor(gather, feats)
                feats.record_stream(comm)
            if host:
                d2h.wait_stream(cur)
                with torch.cuda.stream(d2h):
                    out_h.copy_(feats, non_blocking=True)
                feats.record_stream(d2h)
        if world > 1:
            cur.wait_stream(comm)                # the timed region ends after the last all-gather ...
        if host:
            cur.wait_stream(d2h)                 # ... and the last read-back

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
    lat_sorted = list(step_ms)
    if pipelined and model._pairable((text_d, lengths, noise_d), (text_d, lengths, noise_d)):
        # launches of the pipelined schedule: sample_stream samples two batches per reverse-loop call (two chains in one graph)
        t2, l2, z2 = model._merge_pair((text_d, lengths, noise_d), (text_d, lengths, noise_d))
        zz = model._diffusion_reverse(t2, l2, latents=z2)
        n1 = eng.last_launch_count
        if B * max(lengths) > model.DECODE_MERGE_MIN_ROWS:        # sample_stream then decodes the pair in one call
            model.vae.decode(zz, l2)
            launches = (n1 + eng.last_launch_count) / 2.0
        else:
            model.vae.decode(zz[:, :B].contiguous(), lengths)
            launches = n1 / 2.0 + eng.last_launch_count
        del zz, t2, z2
        torch.cuda.synchronize()
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
        # the dominant kernel of the step (43 % of the launch-list time, profiles/r02b_launches.summary.txt): the fused
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
        # BASELINE config 4: KIT-ML (251-d features, 196 frames), batch 256, bf16 path (2560 latent rows > 1776: the reverse loop runs the separate fused linears, not k_ffn_swap)
        kcfg = L.default_config("kit", num_inference_timesteps=STEPS_DDIM)
        torch.manual_seed(1234)
        kit = LADIFF(kcfg, SyntheticDataModule(251, 21)).to(dev).eval()
        kit.text_encoder = None
        Bk = 256
        tk = torch.randn((2 * Bk, 1, 768), generator=g).to(dev)
        nk = torch.randn((Bk, 5, 256), generator=g).to(dev)
        lk = [FRAMES] * Bk
        for m in ("bf16", "bf16x3"):
            kit.set_precision(m)
            ms_k2 = t_ms(lambda: kit.vae.decode(kit._diffusion_reverse(tk, lk, latents=nk), lk), 3)
            extra[f"kit256_{m}"] = {"value": Bk / (ms_k2 / 1e3), "unit": "seq/s", "ms": ms_k2,
                                    "config": "KIT-ML 251-d features, 196 frames, batch 256, 50-step DDIM + CFG 7.5 + decode"}
        del kit
        # CPU baseline: the reference path (oracle port) on this box's host cores, on THIS model's weights and THIS step's inputs --
        # one complete batch (all 50 CFG steps + decode, nothing extrapolated); its output doubles as the live parity check of
        # the headline configuration in both tensor-core modes
        from oracle import ladiff_oracle as O
        try:
            torch.set_num_threads(len(os.sched_getaffinity(0)))     # all host threads, whatever OMP_NUM_THREADS the launcher exported
        except Exception:
            pass
        sd = {"denoiser." + k: v.detach().float().cpu() for k, v in model.denoiser.state_dict().items()}
        sd.update({"vae." + k: v.detach().float().cpu() for k, v in model.vae.state_dict().items()})
        cpu_reference_run(O, sd, text_h, noise_h, lengths, 2)
        sec, ref_feats = cpu_reference_run(O, sd, text_h, noise_h, lengths)
        cpu = {"value": B / sec, "unit": "seq/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"B={B} L={FRAMES}: one complete batch, all {STEPS_DDIM} CFG denoiser steps + DDIM + decode, fp32 torch CPU (oracle/ladiff_oracle.py)"}
        parity = {"ref_abs_max": float(ref_feats.abs().max()), "against": "cpu_baseline output on the same weights / inputs (oracle as checker)"}
        for m in ("bf16x3", "bf16"):
            model.set_precision(m)
            f = model.vae.decode(model._diffusion_reverse(text_d, lengths, latents=noise_d), lengths).cpu()
            parity[f"{m}_max_abs_err"] = float((f - ref_feats).abs().max())
        model.set_precision(args.mode)
        extra["parity"] = parity
        # BASELINE config 1: batch 1, 196 frames, 50 steps on the CPU (reference path), and the same on the GPU
        t1, n1_, l1 = text_h[[0, B]], noise_h[:1], [FRAMES]
        cpu_reference_run(O, sd, t1, n1_, l1, 2)
        sec1, _ = cpu_reference_run(O, sd, t1, n1_, l1)
        ms_g1 = t_ms(lambda: model.vae.decode(model._diffusion_reverse(t1.to(dev), l1, latents=n1_.to(dev)), l1), 5)
        extra["batch1"] = {"cpu_latency_ms": sec1 * 1e3, "cpu_seq_s": 1.0 / sec1, "cores": torch.get_num_threads(),
                           "gpu_latency_ms": ms_g1, "gpu_seq_s": 1e3 / ms_g1,
                           "config": "batch 1, 196 frames, 50-step DDIM + CFG 7.5 + decode (BASELINE config 1)"}
    else:
        cpu = None

    out = {
        "metric": METRIC, "value": value, "unit": "seq/s", "n_gpus": world,
        "steps": args.steps, "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None,
        "dtype": {"bf16x3": "f16 hi/lo split x3 products, fp32 accumulate (fp32-grade, parity 1e-3)", "bf16": "bf16, fp32 accumulate",
                  "fp32": "f32"}[args.mode],
        "data": "synthetic",
        "config": workload_config(B),
        "run": {"mode": args.mode, "l2": "L2 flushed (256 MiB write) between timed iterations",
                "collective": "all_gather_into_tensor of motions per step" if world > 1 else "none",
                "schedule": ("pipelined over the K steps (LADIFF.sample_stream: two consecutive batches of 128 share one reverse-loop launch "
                             "as two independent chains of one CUDA graph, the pair is decoded in one call on a low-priority stream under the "
                             "reverse loop of the next pair; results bit-identical to one batch at a time -- latency_ms_per_batch / p50 are "
                             "the one-batch-at-a-time figures)") if pipelined else "one batch at a time"},
        "latency_ms_per_batch": ms_seq / args.steps,
        "p50_latency_ms": lat_sorted[len(lat_sorted) // 2], "p90_latency_ms": lat_sorted[min(len(lat_sorted) - 1, int(0.9 * len(lat_sorted)))], "value_sequential": world * B * args.steps / (ms_seq / 1e3),
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
                     "of the ncu --set full capture in profiles/r02_ncu_full_kernels.txt"),
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
    ap.add_argument("--sweep", type=int, default=0, help="BASELINE config 5: this many prompts sharded over the ranks (strong scaling)")
    ap.add_argument("--micro", type=int, default=1024, help="micro-batch (prompts per call) of the sweep; 1024 measured best on one B200 (7.6 k seq/s against 5.9 k at 128 and 5.5 k at 256)")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    elif args.sweep > 0:
        run_sweep(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
