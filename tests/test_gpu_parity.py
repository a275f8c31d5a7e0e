"""Parity of the CUDA path (through the C ABI) with the reference's own outputs (tests/golden, produced by the
unmodified reference modules) and with the CPU oracle on seeded inputs."""
import os

import numpy as np
import pytest
import torch

from oracle import ladiff_oracle as O

pytestmark = pytest.mark.gpu

# max-abs tolerance on decoded features.  fp32 / x3 ("bf16x3" in the API: fp16 hi/lo split operands, 3 products): the 1e-3 contract of
# north_star.  bf16: 2 x the largest error MEASURED on the stated configurations (B = 128 x 196 frames: 0.139; ragged: 0.144;
# KIT B = 256: 0.147 -- the `[bf16 measured]` lines of this file, also printed live by bench.py as parity.bf16_max_abs_err) on
# features with abs-max 2.5 - 4.4: single-pass bf16 operands (8 mantissa bits) under the ~100x amplification of the 50-step CFG loop.
FEATS_TOL = {"fp32": 1e-3, "bf16x3": 1e-3, "bf16": 0.30}


def ddim_tables(n):
    acp = O.ddim_alphas_cumprod().double()
    ts = O.ddim_timesteps(n)
    c1, c2 = [], []
    for t in ts:
        p = int(t) - 1000 // n
        a_t = acp[int(t)]
        a_p = acp[p] if p >= 0 else acp[0]
        c1.append(float((a_p / a_t).sqrt()))
        c2.append(float((1 - a_p).sqrt() - a_p.sqrt() * (1 - a_t).sqrt() / a_t.sqrt()))
    return [int(t) for t in ts], c1, c2


@pytest.mark.parametrize("mode", ["fp32", "bf16x3", "bf16"])
def test_denoiser_forward_vs_reference(engine, golden_dir, mode):
    from ladiff_b200._lib import MODES
    G = np.load(os.path.join(golden_dir, "denoiser_step.npz"))
    lengths = G["lengths"].tolist()
    B = len(lengths)
    g = torch.Generator().manual_seed(int(G["input_seed"]))
    text = torch.randn((2 * B, 1, 768), generator=g)
    x = O.initial_latents(torch.randn((B, 5, 256), generator=g), lengths)
    mie = O.max_iter_elements_of(lengths).tolist() * 2
    out = engine.denoiser_forward(torch.cat([x] * 2).cuda(), int(G["timestep"]), text.cuda(), mie, MODES[mode]).cpu()
    ref = torch.from_numpy(G["out"])
    valid = O.latent_mask_of(torch.tensor(mie))
    err = (out - ref)[valid].abs().max().item()
    assert (out[~valid] == 0).all()
    tol = {"fp32": 5e-5, "bf16x3": 2e-4, "bf16": 0.1}[mode]
    assert err < tol, f"{mode}: denoiser.forward max-abs err {err:.3e} (ref scale {ref.abs().max():.2f})"


@pytest.mark.parametrize("mode", ["fp32", "bf16x3", "bf16"])
def test_sampling_vs_reference(engine, golden_dir, mode):
    from ladiff_b200._lib import MODES
    G = np.load(os.path.join(golden_dir, "sampling.npz"))
    lengths = G["lengths"].tolist()
    text, noise, _ = O.synthetic_inputs(len(lengths), seed=int(G["input_seed"]))
    ts, c1, c2 = ddim_tables(50)
    z = engine.diffusion_reverse(text.cuda(), lengths, noise.cuda(), ts, c1, c2, 7.5, MODES[mode])
    feats = engine.vae_decode(z, lengths, MODES[mode]).cpu()
    z = z.cpu()
    zref, fref = torch.from_numpy(G["z"]), torch.from_numpy(G["feats"])
    m = O.max_iter_elements_of(lengths)
    for b, mb in enumerate(m):
        assert (z[int(mb):, b] == 0).all(), "masked latent rows must be exact zeros (ladiff.py:562-566)"
    for b, L in enumerate(lengths):
        assert (feats[b, L:] == 0).all(), "padded frames must be exact zeros (ladiff_vae.py:358)"
    zerr, ferr = (z - zref).abs().max().item(), (feats - fref).abs().max().item()
    print(f"[{mode}] latents max-abs err {zerr:.3e} (scale {zref.abs().max():.1f}); feats max-abs err {ferr:.3e}")
    assert ferr < FEATS_TOL[mode], f"{mode}: decoded features max-abs err {ferr:.3e}"
    # 20-step schedule (the shipped YAML default, configs/modules/scheduler.yaml:3)
    ts, c1, c2 = ddim_tables(20)
    z20 = engine.diffusion_reverse(text.cuda(), lengths, noise.cuda(), ts, c1, c2, 7.5, MODES[mode]).cpu()
    e20 = (z20 - torch.from_numpy(G["z20"])).abs().max().item()
    print(f"[{mode}] 20-step latents max-abs err {e20:.3e} (scale {torch.from_numpy(G['z20']).abs().max():.1f})")
    assert e20 < {"fp32": 5e-3, "bf16x3": 2e-2, "bf16": 50.0}[mode], f"{mode}: 20-step latents err {e20:.3e}"


@pytest.mark.parametrize("mode", ["fp32", "bf16x3", "bf16"])
@pytest.mark.parametrize("name", ["decode", "decode_kit"])
def test_decode_vs_reference(engine, golden_dir, oracle_sd, mode, name):
    from ladiff_b200._lib import MODES, Engine
    G = np.load(os.path.join(golden_dir, name + ".npz"))
    lengths = G["lengths"].tolist()
    fref = torch.from_numpy(G["feats"])
    nf = fref.shape[-1]
    eng = engine
    if nf != 263:
        sdk = O.make_state_dict(1234, nf, perturb=True)
        eng = Engine(nfeats=nf)
        eng.set_weights({k: v.cuda() for k, v in O.sub(sdk, "vae.").items()}, "vae.")
        eng.finalize(2)
    g = torch.Generator().manual_seed(int(G["input_seed"]))
    zin = O.initial_latents(torch.randn((len(lengths), 5, 256), generator=g), lengths).permute(1, 0, 2).contiguous()
    feats = eng.vae_decode(zin.cuda(), lengths, MODES[mode]).cpu()
    assert feats.shape == fref.shape
    for b, L in enumerate(lengths):
        assert (feats[b, L:] == 0).all()
    err = (feats - fref).abs().max().item()
    tol = {"fp32": 1e-4, "bf16x3": 1e-3, "bf16": 0.25}[mode]
    print(f"[{mode}] {name}: feats max-abs err {err:.3e}")
    assert err < tol, f"{mode} {name}: max-abs err {err:.3e}"


def test_cfg_ddim_step(engine):
    g = torch.Generator().manual_seed(3)
    B = 37
    pred = torch.randn((2 * B, 5, 256), generator=g)
    lat = torch.randn((B, 5, 256), generator=g)
    out = engine.cfg_ddim_step(pred.cuda(), lat.cuda(), 7.5, 1.12285173, -0.12325197).cpu()
    u, c = pred.chunk(2)
    eps = u + 7.5 * (c - u)
    acp = O.ddim_alphas_cumprod()
    ref = O.ddim_step(eps, 981, lat, acp, 50)
    assert (out - ref).abs().max().item() < 2e-5


def test_feats2joints(engine):
    g = torch.Generator().manual_seed(5)
    B, L = 6, 196
    feats = torch.randn((B, L, 263), generator=g) * 0.5
    mean, std = torch.randn((263,), generator=g) * 0.1, torch.rand((263,), generator=g) + 0.5
    out = engine.feats2joints(feats.cuda(), mean.cuda(), std.cuda(), 22).cpu()
    ref = O.feats2joints(feats, mean, std, 22)
    err = (out - ref).abs().max().item()
    assert out.shape == ref.shape and err < 5e-3, f"feats2joints max-abs err {err:.3e} (scale {ref.abs().max():.1f})"


@pytest.mark.parametrize("mode", ["fp32", "bf16x3"])
def test_batch32_ragged_vs_oracle(engine, oracle_sd, mode):
    """Seeded ragged batch (SURVEY.md 8d lengths), CUDA path vs the CPU oracle run here."""
    from ladiff_b200._lib import MODES
    B = 32
    text, noise, lengths = O.synthetic_inputs(B, seed=1234, ragged=True)
    ts, c1, c2 = ddim_tables(50)
    z = engine.diffusion_reverse(text.cuda(), lengths, noise.cuda(), ts, c1, c2, 7.5, MODES[mode])
    feats = engine.vae_decode(z, lengths, MODES[mode]).cpu()
    ref = O.sample_motion(oracle_sd, text, lengths, noise)
    err = (feats - ref).abs().max().item()
    print(f"[{mode}] B=32 ragged feats max-abs err {err:.3e}")
    assert err < 1e-3


@pytest.mark.parametrize("mode", ["bf16x3", "bf16"])
def test_batch128_properties(engine, mode):
    """Full-size config (B=128, 196 frames): batch-composition independence -- a sample decoded inside the batch
    equals the same sample run alone (attention never crosses sequences), and repeat runs are bit-identical."""
    from ladiff_b200._lib import MODES
    B = 128
    text, noise, lengths = O.synthetic_inputs(B, seed=99, ragged=False)
    ts, c1, c2 = ddim_tables(50)
    z = engine.diffusion_reverse(text.cuda(), lengths, noise.cuda(), ts, c1, c2, 7.5, MODES[mode])
    f1 = engine.vae_decode(z, lengths, MODES[mode])
    z2 = engine.diffusion_reverse(text.cuda(), lengths, noise.cuda(), ts, c1, c2, 7.5, MODES[mode])
    assert torch.equal(z, z2), "graph replay must be deterministic"
    idx = [0, 77, 127]
    sub_text = torch.cat([text[idx], text[[B + i for i in idx]]])
    zs = engine.diffusion_reverse(sub_text.cuda(), [lengths[i] for i in idx], noise[idx].cuda(), ts, c1, c2, 7.5, MODES[mode])
    fs = engine.vae_decode(zs, [lengths[i] for i in idx], MODES[mode])
    assert torch.isfinite(f1).all()
    # different batch sizes may pick different tilings (cluster-split vs whole-row LayerNorm epilogue), i.e. a different
    # fp32 summation order: equal up to rounding, not bit-equal
    tol = {"bf16x3": 2e-3, "bf16": 0.25}[mode]
    err = (fs - f1[idx]).abs().max().item()
    assert err < tol, f"a sample's result must not depend on its batch neighbours (max-abs diff {err:.3e})"


@pytest.mark.parametrize("mode", ["bf16x3", "bf16"])
@pytest.mark.parametrize("ragged", [False, True])
def test_denoiser_forward_large_batch_vs_oracle(engine, oracle_sd, mode, ragged, monkeypatch):
    """One denoiser.forward at the batch of the headline config (token groups of 48, the sa_block attention fused into the
    feed-forward kernel), full-length and ragged (sequences of 1..5 latent rows, so the 12 rows a CTA owns span several
    sequences and the last groups are partly / wholly empty), against the CPU oracle and against the unfused kernels."""
    from ladiff_b200._lib import MODES
    monkeypatch.setenv("LADIFF_ATT_FUSE", "1")     # opt-in path (measured slower than the separate launch)
    B = 128      # 2 * 128 * 5 = 1280 row slots -> token groups of 48 whatever the lengths are (the plan is sized for the maximum)
    text, noise, lengths = O.synthetic_inputs(B, seed=4321, ragged=ragged)
    x = O.initial_latents(noise, lengths)
    mie = O.max_iter_elements_of(lengths).tolist() * 2
    out = engine.denoiser_forward(torch.cat([x] * 2).cuda(), 481, text.cuda(), mie, MODES[mode]).cpu()
    ref = O.denoiser_forward(oracle_sd, torch.cat([x] * 2), torch.tensor(481), text, torch.tensor(mie))
    valid = O.latent_mask_of(torch.tensor(mie))
    err = (out - ref)[valid].abs().max().item()
    tol = {"bf16x3": 2e-4, "bf16": 0.1}[mode]
    assert (out[~valid] == 0).all()
    assert err < tol, f"{mode} ragged={ragged}: fused path max-abs err {err:.3e}"
    monkeypatch.delenv("LADIFF_ATT_FUSE")
    out2 = engine.denoiser_forward(torch.cat([x] * 2).cuda(), 481, text.cuda(), mie, MODES[mode]).cpu()
    err2 = (out2 - out)[valid].abs().max().item()
    assert err2 < (1e-4 if mode == "bf16x3" else 0.1), f"{mode}: fused vs unfused attention differ by {err2:.3e}"


def test_sample_stream_matches_sequential(oracle_sd):
    """LADIFF.sample_stream (decode of batch i overlapped with the reverse loop of batch i+1 on two streams) returns exactly
    what the one-batch-at-a-time path returns, in order, for batches of different sizes / lengths."""
    import ladiff_b200 as L
    from ladiff_b200.data import SyntheticDataModule
    from ladiff_b200.modeltype import LADIFF
    torch.set_grad_enabled(False)
    model = LADIFF(L.default_config("humanml3d", num_inference_timesteps=6), SyntheticDataModule(263, 22))
    model.denoiser.load_state_dict(O.sub(oracle_sd, "denoiser."), strict=True)
    model.vae.load_state_dict(O.sub(oracle_sd, "vae."), strict=True)
    model = model.to("cuda:0").eval()
    g = torch.Generator().manual_seed(77)
    batches = []
    for B, lengths in ((3, [196, 52, 120]), (5, [40, 44, 196, 100, 148]), (3, [196, 52, 120]), (3, [64, 64, 64])):
        batches.append((torch.randn((2 * B, 1, 768), generator=g).cuda(), lengths, torch.randn((B, 5, 256), generator=g).cuda()))
    seq = [model.sample_features(t, l, latents=n).clone() for t, l, n in batches]
    torch.cuda.synchronize()
    out = [f.clone() for f in model.sample_stream(iter(batches))]
    torch.cuda.synchronize()
    assert len(out) == len(seq)
    for a, b in zip(seq, out):
        assert a.shape == b.shape and torch.equal(a, b)


def test_sample_stream_pairs_batches(oracle_sd):
    """Batches that fit are sampled two per reverse-loop launch (two chains in one graph): still bit-identical, batch by batch, to
    the one-batch-at-a-time path -- ragged lengths, batches that cannot be paired (unequal sizes) in between, and pair=False."""
    import ladiff_b200 as L
    from ladiff_b200.data import SyntheticDataModule
    from ladiff_b200.modeltype import LADIFF
    torch.set_grad_enabled(False)
    model = LADIFF(L.default_config("humanml3d", num_inference_timesteps=4), SyntheticDataModule(263, 22))
    model.denoiser.load_state_dict(O.sub(oracle_sd, "denoiser."), strict=True)
    model.vae.load_state_dict(O.sub(oracle_sd, "vae."), strict=True)
    model = model.to("cuda:0").eval()
    g = torch.Generator().manual_seed(78)
    batches = []
    # pairs: (80, 80) [80 x 120 frames < 75 row tiles: two decodes], the two 90s [one of them small: two decodes], and the two
    # 100s [both past 74 row tiles, different max lengths: ONE merged decode, sliced]
    for B, hi in ((80, 196), (80, 120), (96, 196), (70, 196), (90, 64), (90, 196), (90, 100), (100, 196), (100, 150)):
        lengths = [int(x) for x in torch.randint(20, hi + 1, (B,), generator=g)]
        lengths[0] = hi
        batches.append((torch.randn((2 * B, 1, 768), generator=g).cuda(), lengths, torch.randn((B, 5, 256), generator=g).cuda()))
    assert model._pairable(batches[0], batches[1]) and not model._pairable(batches[2], batches[3]) and model._pairable(batches[5], batches[6])
    assert model._pairable(batches[7], batches[8]) and all(len(l) * max(l) > model.DECODE_MERGE_MIN_ROWS for _, l, _ in batches[7:9])
    seq = [model.sample_features(t, l, latents=n).clone() for t, l, n in batches]
    torch.cuda.synchronize()
    for pair in (True, False):
        out = [f.clone() for f in model.sample_stream(iter(batches), pair=pair)]
        torch.cuda.synchronize()
        assert len(out) == len(seq)
        for a, b in zip(seq, out):
            assert a.shape == b.shape and torch.equal(a, b), f"pair={pair}"


# ------------------------------------------------------------------------------------------------------------------
# round 2: the stated configurations themselves against the oracle (VERDICT r1 "parity partial")
def _bf16_report(name, err):
    """bf16 mode: the measured error is recorded next to the bound so that the stated tolerance stays 'measured'."""
    print(f"[bf16 measured] {name}: max-abs err {err:.4e}")


@pytest.mark.parametrize("mode", ["fp32", "bf16x3", "bf16"])
@pytest.mark.parametrize("ragged", [False, True])
def test_headline_batch128_vs_oracle(engine, oracle_sd, mode, ragged):
    """BASELINE config 3 itself: B = 128, 50-step DDIM + CFG 7.5 + decode (196 frames, and the seeded ragged batch) against
    ``O.sample_motion``.  At 25 088 frame rows the decoder runs the whole-row LayerNorm-epilogue GEMM (not the cluster-split
    one the small cases take) and the 784-CTA decoder GEMMs; the reverse loop runs token groups of 48."""
    from ladiff_b200._lib import MODES
    B = 128
    text, noise, lengths = O.synthetic_inputs(B, seed=2024, ragged=ragged)
    ts, c1, c2 = ddim_tables(50)
    z = engine.diffusion_reverse(text.cuda(), lengths, noise.cuda(), ts, c1, c2, 7.5, MODES[mode])
    feats = engine.vae_decode(z, lengths, MODES[mode]).cpu()
    zref = O.diffusion_reverse(oracle_sd, text, lengths, noise, 50, 7.5)
    ref = O.vae_decode(oracle_sd, zref, lengths)
    for b, L in enumerate(lengths):
        assert (feats[b, L:] == 0).all()
    err = (feats - ref).abs().max().item()
    zerr = (z.cpu() - zref).abs().max().item()
    print(f"[{mode}] B=128 ragged={ragged}: feats max-abs err {err:.3e} (scale {ref.abs().max():.2f}); latents err {zerr:.3e} (scale {zref.abs().max():.1f})")
    if mode == "bf16":
        _bf16_report(f"B=128 ragged={ragged} feats", err)
    assert err < FEATS_TOL[mode], f"{mode} ragged={ragged}: decoded features max-abs err {err:.3e}"


@pytest.mark.parametrize("mode", ["bf16x3", "bf16"])
def test_kit_batch256_vs_oracle(mode):
    """BASELINE config 4: KIT-ML (251-d features), batch 256, full sampling.  2560 latent rows > 1776: the reverse loop takes
    the four separate fused linears of the feed-forward pairs instead of k_ffn_swap, a path the HumanML3D headline never runs."""
    from ladiff_b200._lib import MODES, Engine
    sdk = O.make_state_dict(1234, 251, perturb=True)
    eng = Engine(nfeats=251)
    eng.set_weights({k: v.cuda() for k, v in O.sub(sdk, "denoiser.").items()}, "denoiser.")
    eng.set_weights({k: v.cuda() for k, v in O.sub(sdk, "vae.").items()}, "vae.")
    eng.finalize(3)
    B = 256
    text, noise, lengths = O.synthetic_inputs(B, seed=251, ragged=True)
    lengths = [196 if i % 3 == 0 else L for i, L in enumerate(lengths)]      # a third at the full 196 frames
    ts, c1, c2 = ddim_tables(50)
    z = eng.diffusion_reverse(text.cuda(), lengths, noise.cuda(), ts, c1, c2, 7.5, MODES[mode])
    feats = eng.vae_decode(z, lengths, MODES[mode]).cpu()
    ref = O.sample_motion(sdk, text, lengths, noise)
    assert feats.shape == ref.shape == (B, 196, 251)
    err = (feats - ref).abs().max().item()
    print(f"[{mode}] KIT B=256: feats max-abs err {err:.3e} (scale {ref.abs().max():.2f})")
    if mode == "bf16":
        _bf16_report("KIT B=256 feats", err)
    assert err < FEATS_TOL[mode]


def _model(oracle_sd, n_steps=6, **cfg_over):
    import ladiff_b200 as L
    from ladiff_b200.data import SyntheticDataModule
    from ladiff_b200.modeltype import LADIFF
    torch.set_grad_enabled(False)
    cfg = L.default_config("humanml3d", num_inference_timesteps=n_steps)
    for k, v in cfg_over.items():
        cfg[k] = v
    g = torch.Generator().manual_seed(9)
    mean, std = 0.1 * torch.randn((263,), generator=g), 0.5 + torch.rand((263,), generator=g)
    model = LADIFF(cfg, SyntheticDataModule(263, 22, mean=mean, std=std))
    model.denoiser.load_state_dict(O.sub(oracle_sd, "denoiser."), strict=True)
    model.vae.load_state_dict(O.sub(oracle_sd, "vae."), strict=True)
    return model.to("cuda:0").eval(), mean, std


def test_forward_with_strings_and_gen_from_latent(oracle_sd):
    """SURVEY 8 rows a1 / a15 through the mirror classes: ``model({"text": [...], "length": [...]})`` (CLIP -> CFG text build ->
    reverse loop -> decode -> feats2joints -> remove_padding) and ``gen_from_latent`` against the oracle on the same CLIP
    embeddings / initial noise."""
    model, mean, std = _model(oracle_sd)
    texts = ["a person walks forward", "someone jumps twice", "a person walks forward"]
    lengths = [196, 52, 120]
    torch.manual_seed(77)                                     # the reference draws the initial noise from the global RNG (:380-385)
    joints = model({"text": texts, "length": lengths})
    assert [tuple(j.shape) for j in joints] == [(L, 22, 3) for L in lengths] and all(not j.is_cuda for j in joints)
    # the same thing step by step with the oracle
    emb = model.text_encoder([""] * 3 + texts)                # uncond first (ladiff.py:258-264)
    assert emb.shape == (6, 1, 768) and torch.equal(emb[3], emb[5])
    torch.manual_seed(77)
    noise = torch.randn((3, 5, 256), device="cuda", dtype=torch.float)
    zref = O.diffusion_reverse(oracle_sd, emb.cpu(), lengths, noise.cpu(), 6, 7.5)
    fref = O.vae_decode(oracle_sd, zref, lengths)
    jref = O.feats2joints(fref, mean, std, 22)
    # (i) the decoded features behind forward() (same RNG draw again) against the oracle's
    torch.manual_seed(77)
    feats = model.sample_features(emb, lengths).cpu()
    ferr = (feats - fref).abs().max().item()
    assert ferr < 1e-3, f"forward(): decoded features max-abs err {ferr:.3e}"
    # (ii) the joints forward() returned against recover_from_ric of exactly those features.  (Against jref the root rotation
    # integrates the 1e-5 feature differences over up to 196 frames and multiplies them by the O(30) positions: not a parity
    # measure of the kernels.)
    jmine = O.feats2joints(feats, mean, std, 22)
    for j, L, r, r0 in zip(joints, lengths, jmine, jref):
        err, scale = (j - r[:L]).abs().max().item(), r[:L].abs().max().item()
        print(f"forward(): joints max-abs err {err:.3e} (scale {scale:.1f}); vs oracle-from-text {(j - r0[:L]).abs().max().item():.3e}")
        assert err < 2e-4 * max(1.0, scale), f"forward(): joints max-abs err {err:.3e} (scale {scale:.1f})"
        assert (j - r0[:L]).abs().max().item() < 5e-3 * max(1.0, scale)
    # gen_from_latent (ladiff.py:310-318): decode-only entry
    out = model.gen_from_latent({"latent": zref.cuda(), "length": lengths})
    for j, L, r in zip(out, lengths, jref):
        assert j.shape == (L, 22, 3) and (j - r[:L]).abs().max().item() < 5e-3 * max(1.0, r[:L].abs().max().item())


def test_ddpm_sampling_vs_oracle(engine, oracle_sd):
    """DDPM ancestral sampling (x' = c1 x + c2 eps + c3 noise) with the variance noise injected, against the
    oracle's DDPMScheduler.step restatement; then the in-kernel Philox stream: deterministic per seed, different across seeds."""
    from ladiff_b200._lib import MODE_BF16X3
    from ladiff_b200.scheduler import DDPMScheduler
    s = DDPMScheduler(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                      variance_type="fixed_small", clip_sample=False)
    n = 20
    s.set_timesteps(n)
    ts, c1, c2, c3 = s.fused_coefficients()
    lengths = [196, 40, 100, 148]
    text, noise, _ = O.synthetic_inputs(len(lengths), seed=31)
    step_noise = torch.randn((n, len(lengths), 5, 256), generator=torch.Generator().manual_seed(32))
    z = engine.diffusion_reverse(text.cuda(), lengths, noise.cuda(), ts, c1, c2, 7.5, MODE_BF16X3, c3=c3, step_noise=step_noise.cuda())
    zref = O.diffusion_reverse(oracle_sd, text, lengths, noise, n, 7.5, scheduler="ddpm", step_noise=step_noise)
    err = (z.cpu() - zref).abs().max().item()
    print(f"DDPM latents max-abs err {err:.3e} (scale {zref.abs().max():.1f})")
    assert err < 2e-4 * zref.abs().max().item(), f"DDPM latents max-abs err {err:.3e} (scale {zref.abs().max():.1f})"
    f = engine.vae_decode(z, lengths, MODE_BF16X3).cpu()
    ferr = (f - O.vae_decode(oracle_sd, zref, lengths)).abs().max().item()
    assert ferr < 2e-3, f"DDPM decoded features max-abs err {ferr:.3e}"
    za = engine.diffusion_reverse(text.cuda(), lengths, noise.cuda(), ts, c1, c2, 7.5, MODE_BF16X3, c3=c3, seed=5)
    zb = engine.diffusion_reverse(text.cuda(), lengths, noise.cuda(), ts, c1, c2, 7.5, MODE_BF16X3, c3=c3, seed=5)
    zc = engine.diffusion_reverse(text.cuda(), lengths, noise.cuda(), ts, c1, c2, 7.5, MODE_BF16X3, c3=c3, seed=6)
    assert torch.equal(za, zb) and not torch.equal(za, zc) and torch.isfinite(za).all()
    # Philox + Box-Muller statistics: with c1 = c2 = 0, c3 = 1 the loop's state after a step IS the drawn noise (last step has t = 0 -> no noise,
    # so use n - 1 noisy steps and read the latents through a zero-coefficient final step)
    zero, one = [0.0] * n, [1.0] * (n - 1) + [0.0]
    keep = [0.0] * (n - 1) + [1.0]
    zn = engine.diffusion_reverse(text.cuda(), [196] * 4, noise.cuda(), ts, keep, zero, 7.5, MODE_BF16X3, c3=one, seed=11).cpu()
    assert abs(zn.mean().item()) < 0.05 and abs(zn.std().item() - 1.0) < 0.05


def test_ardiff_branch_vs_oracle(oracle_sd):
    """the ARDIFF autoregressive branch through LADIFF._diffusion_reverse (both motion_conditioning modes)
    against the oracle restatement of ladiff.py:419-467."""
    for mc in ("last", "full"):
        model, _, _ = _model(oracle_sd, n_steps=5, ARDIFF=True)
        model.motion_conditioning = mc
        lengths = [196, 100, 52]
        text, noise, _ = O.synthetic_inputs(3, seed=41)
        z = model._diffusion_reverse(text.cuda(), lengths, latents=noise.cuda()).cpu()
        zref = O.diffusion_reverse_ardiff(oracle_sd, text, lengths, noise, 5, 7.5, motion_conditioning=mc)
        assert z.shape == zref.shape == (5, 3, 256)
        err = (z - zref).abs().max().item()
        assert (z[3:, 1] == 0).all() and (z[2:, 2] == 0).all()
        print(f"ARDIFF ({mc}) latents max-abs err {err:.3e} (scale {zref.abs().max():.1f})")
        assert err < 2e-4 * zref.abs().max().item(), f"ARDIFF ({mc}) latents max-abs err {err:.3e} (scale {zref.abs().max():.1f})"


@pytest.mark.parametrize("mode", ["fp32", "bf16x3", "bf16"])
def test_vae_encode_vs_reference(engine, golden_dir, mode):
    """LADiffVae.encode on the CUDA path (ragged mu | logvar | frames tokens through the non-MD skip encoder)
    against dist.loc / dist.scale of the unmodified reference module (tests/golden/encode.npz)."""
    from ladiff_b200._lib import MODES
    G = np.load(os.path.join(golden_dir, "encode.npz"))
    lengths = G["lengths"].tolist()
    g = torch.Generator().manual_seed(int(G["input_seed"]))
    motion = 0.5 * torch.randn((len(lengths), max(lengths), 263), generator=g)
    for i, L in enumerate(lengths):
        motion[i, L:] = 0
    eps = torch.randn((5, len(lengths), 256), generator=torch.Generator().manual_seed(1))
    lat, mu, std = (t.cpu() for t in engine.vae_encode(motion.cuda(), lengths, MODES[mode], eps.cuda()))
    valid = O.latent_mask_of(torch.from_numpy(G["mie"])).T
    mref, sref = torch.from_numpy(G["mu"]), torch.from_numpy(G["std"])
    merr = (mu - mref)[valid].abs().max().item()
    serr = ((std - sref)[valid].abs() / sref[valid]).max().item()
    print(f"[{mode}] encode: mu max-abs err {merr:.3e} (scale {mref[valid].abs().max():.2f}), std max-rel err {serr:.3e}")
    tol = {"fp32": 1e-4, "bf16x3": 1e-3, "bf16": 0.25}[mode]
    assert merr < tol and serr < tol
    assert (lat[~valid] == 0).all() and torch.allclose(lat[valid], (mu + std * eps)[valid], atol=1e-6)


def test_t2m_eval_glue(oracle_sd):
    """LADIFF.t2m_eval (ladiff.py:1111-1282) in both stages: every key of the reference's result dict, rows aligned by
    descending length, reconstruction = decode(encode(motion)) resp. decode(sample) checked against the oracle."""
    B, lengths = 4, [100, 196, 52, 148]
    g = torch.Generator().manual_seed(3)
    motion = 0.5 * torch.randn((B, 196, 263), generator=g)
    for i, L in enumerate(lengths):
        motion[i, L:] = 0
    batch = {"text": [f"motion {i}" for i in range(B)], "motion": motion.cuda(), "length": lengths,
             "word_embs": torch.randn((B, 22, 300), generator=g).cuda(), "pos_ohot": torch.randn((B, 22, 15), generator=g).cuda(),
             "text_len": torch.tensor([22, 20, 9, 5])}
    order = [1, 3, 0, 2]
    for stage in ("vae", "diffusion"):
        model, mean, std = _model(oracle_sd)
        model.stage = stage
        eps = torch.randn((5, B, 256), generator=torch.Generator().manual_seed(8))
        noise = torch.randn((B, 5, 256), generator=torch.Generator().manual_seed(9))
        rs = model.t2m_eval(batch, latents=noise.cuda(), eps=eps.cuda())
        assert set(rs) == {"m_ref", "m_rst", "lat_t", "lat_m", "lat_rm", "joints_ref", "joints_rst"}
        assert rs["m_rst"].shape == rs["m_ref"].shape == (B, 196, 263)
        assert rs["lat_t"].shape == rs["lat_m"].shape == rs["lat_rm"].shape == (B, 512)
        assert rs["joints_rst"].shape == rs["joints_ref"].shape == (B, 196, 22, 3)
        if stage == "vae":
            z, _, _, _ = O.vae_encode(oracle_sd, motion, lengths, eps)
        else:
            emb = model.text_encoder([""] * B + batch["text"]).cpu()
            z = O.diffusion_reverse(oracle_sd, emb, lengths, noise, 6, 7.5)
        fref = O.vae_decode(oracle_sd, z, lengths)           # renorm4t2m is the identity for the synthetic datamodule (eval stats = stats)
        err = (rs["m_rst"].cpu() - fref[order]).abs().max().item()
        assert err < 2e-3, f"t2m_eval[{stage}] m_rst max-abs err {err:.3e}"
        assert torch.allclose(rs["m_ref"].cpu(), motion[order], atol=1e-6)
