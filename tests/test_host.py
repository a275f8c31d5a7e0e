"""CPU: host logic of the drop-in layer -- config loading / retargeting, schedulers, parameter layout, C-ABI exports."""
import ctypes
import os
import re

import pytest
import torch

import ladiff_b200 as L
from ladiff_b200 import build as B
from ladiff_b200.config import Cfg, instantiate_from_config, load_config, merge
from ladiff_b200.scheduler import DDIMScheduler, DDPMScheduler
from oracle import ladiff_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cfg_interpolation_and_merge():
    c = merge({"a": {"b": [7, 256], "c": 1}, "x": "${a.b}", "s": "v=${a.c}"}, {"a": {"c": 2}})
    assert c.x == [7, 256] and c.s == "v=2" and c.a.c == 2
    c.a.c = 5
    assert c.s == "v=5"
    with pytest.raises(KeyError):
        Cfg({"y": "${nope.k}"}).y
    with pytest.raises(KeyError):
        instantiate_from_config(Cfg({"params": {}}))        # reference: KeyError("Expected key `target` ...")


def test_default_configs():
    h = L.default_config("humanml3d")
    k = L.default_config("kit")
    assert (h.DATASET.NFEATS, h.DATASET.NJOINTS, k.DATASET.NFEATS, k.DATASET.NJOINTS) == (263, 22, 251, 21)
    p = h.model.denoiser.params
    assert p.latent_dim == [7, 256] and p.ablation.MAX_IT == 5 and p.ablation.FRAME_PER_LATENT == 48 and p.nfeats == 263
    assert h.model.scheduler.num_inference_timesteps == 50 and h.model.guidance_scale == 7.5
    sch = instantiate_from_config(h.model.scheduler)
    assert isinstance(sch, DDIMScheduler) and not sch.config.clip_sample and sch.config.steps_offset == 1


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir("/root/reference/src/configs"), reason="needs /root/reference")
def test_reference_yaml_files_work_unchanged():
    cfgdir = "/root/reference/src/configs"
    for name, nf in (("config_ladiff_humanml3d.yaml", 263), ("config_ladiff_kit.yaml", 251)):
        c = load_config(os.path.join(cfgdir, name), cfgdir, overrides={"DATASET": {"NFEATS": nf, "NJOINTS": 22, "NCLASSES": 10}})
        assert c.model.denoiser.target == "ladiff_b200.denoiser.LADiffDenoiser"
        assert c.model.motion_vae.target == "ladiff_b200.vae.LADiffVae"
        assert c.model.scheduler.target == "ladiff_b200.scheduler.DDIMScheduler"
        assert c.model.scheduler.num_inference_timesteps == 20          # the shipped YAML value
        den = instantiate_from_config(c.model.denoiser)
        vae = instantiate_from_config(c.model.motion_vae)
        assert den.max_it == 5 and vae.nfeats == nf


def test_ddim_scheduler_matches_oracle_restatement():
    s = DDIMScheduler(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                      clip_sample=False, set_alpha_to_one=False, steps_offset=1)
    assert s.init_noise_sigma == 1.0 and s.config.num_train_timesteps == 1000
    acp = O.ddim_alphas_cumprod()
    assert torch.equal(s.alphas_cumprod, acp)
    for n in (50, 20):
        s.set_timesteps(n)
        assert s.timesteps.dtype == torch.int64 and s.timesteps.tolist() == O.ddim_timesteps(n).tolist()
        ts, c1, c2 = s.fused_coefficients(0.0)
        g = torch.Generator().manual_seed(n)
        x, e = torch.randn((3, 5, 256), generator=g), torch.randn((3, 5, 256), generator=g)
        for i in (0, n // 2, n - 1):
            ref = O.ddim_step(e, ts[i], x, acp, n)
            assert (s.step(e, torch.tensor(ts[i]), x, eta=0.0).prev_sample - ref).abs().max() < 1e-5
            assert (c1[i] * x + c2[i] * e - ref).abs().max() < 1e-5     # the closed form the fused kernel uses
    assert abs(c1[-1] - 1.00042760) < 1e-6 or n != 50
    with pytest.raises(ValueError):
        DDIMScheduler().fused_coefficients()                              # set_timesteps not called
    with pytest.raises(ValueError):
        s.fused_coefficients(eta=0.5)


def test_ddpm_scheduler_basics():
    s = DDPMScheduler(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                      variance_type="fixed_small", clip_sample=False)
    x0, n = torch.ones(2, 3), torch.zeros(2, 3)
    out = s.add_noise(x0, n, torch.tensor([0, 999]))
    assert torch.allclose(out[0], x0[0] * s.alphas_cumprod[0].sqrt()) and torch.allclose(out[1], x0[1] * s.alphas_cumprod[999].sqrt())
    s.set_timesteps(1000)
    assert s.timesteps[0] == 999 and s.timesteps[-1] == 0
    assert s.step(torch.zeros(2, 3), 0, x0).prev_sample.shape == (2, 3)


def test_mirror_modules_have_the_reference_key_layout(oracle_sd):
    from ladiff_b200.denoiser import LADiffDenoiser
    from ladiff_b200.vae import LADiffVae
    abl = L.default_config().TRAIN.ABLATION
    den = LADiffDenoiser(ablation=abl, nfeats=263, latent_dim=[7, 256], num_layers=9)
    vae = LADiffVae(ablation=abl, nfeats=263, latent_dim=[7, 256], arch="encoder_decoder")
    spec = O.state_dict_spec(263)
    assert {"denoiser." + k: tuple(v.shape) for k, v in den.state_dict().items()} == {k: v for k, v in spec.items() if k.startswith("denoiser.")}
    assert {"vae." + k: tuple(v.shape) for k, v in vae.state_dict().items()} == {k: v for k, v in spec.items() if k.startswith("vae.")}
    den.load_state_dict(O.sub(oracle_sd, "denoiser."), strict=True)
    vae.load_state_dict(O.sub(oracle_sd, "vae."), strict=True)
    assert den._dirty and vae._dirty                       # engine weights re-packed on next use
    # in-place parameter updates (optimizer step, EMA, p.data.copy_) are detected too: they move the signature the
    # engine compares before every call (stale packed weights)
    sig = den._param_signature()
    with torch.no_grad():
        den.encoder.norm.weight.mul_(1.0)
    assert den._param_signature() != sig
    # the zero_module-d matrices are re-initialised (xavier) like SkipTransformerEncoder._reset_parameters does
    assert den.encoder.input_blocks[0].ffn.linear2.weight.abs().sum() > 0
    # encode (torch path, outside the CUDA hot path) runs on CPU
    lat, dist, mie = vae.eval().encode(torch.randn(2, 60, 263), [60, 49])
    assert lat.shape == (5, 2, 256) and mie.tolist() == [2, 2] and (lat[2:] == 0).all()


def test_unsupported_configurations_raise_like_the_reference():
    from ladiff_b200.denoiser import LADiffDenoiser
    from ladiff_b200.vae import LADiffVae
    abl = dict(L.default_config().TRAIN.ABLATION.to_dict())
    with pytest.raises(TypeError):
        LADiffDenoiser(ablation=abl, condition="image", latent_dim=[7, 256], num_layers=9)
    with pytest.raises(ValueError):
        LADiffDenoiser(ablation={**abl, "DIFF_PE_TYPE": "xyz"}, latent_dim=[7, 256], num_layers=9)
    with pytest.raises(ValueError):
        LADiffDenoiser(ablation=abl, arch="unet", latent_dim=[7, 256], num_layers=9)
    with pytest.raises(ValueError):
        LADiffVae(ablation={**abl, "PE_TYPE": "xyz"}, nfeats=263, arch="encoder_decoder")
    with pytest.raises(ValueError):
        LADiffVae(ablation=abl, nfeats=263, arch="something")
    with pytest.raises(NotImplementedError):
        LADiffDenoiser(ablation=abl, latent_dim=[7, 512], num_layers=9)


def test_no_cpu_fallback():
    """The product path must fail loudly without the CUDA device / extension (no silent CPU route)."""
    from ladiff_b200.vae import LADiffVae
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    vae = LADiffVae(ablation=L.default_config().TRAIN.ABLATION, nfeats=263, latent_dim=[7, 256], arch="encoder_decoder")
    with pytest.raises(RuntimeError):
        vae.decode(torch.zeros(5, 1, 256), [40])


def test_text_encoder_dedup_and_shape():
    from ladiff_b200.text_encoder import MldTextEncoder, SYNTHETIC
    enc = MldTextEncoder(SYNTHETIC).eval()
    e = enc(["", "a person walks", "", "a person walks"])
    assert e.shape == (4, 1, 768) and torch.equal(e[0], e[2]) and torch.equal(e[1], e[3]) and not torch.equal(e[0], e[1])
    e2 = enc(["", "somebody jumps"])                          # cached "" embedding
    assert torch.equal(e2[0], e[0])
    with pytest.raises(FileNotFoundError):
        MldTextEncoder("/nonexistent/clip-vit-large-patch14")
    # the cached "" embedding must not survive a weight (re)load, and is never used in training / finetune mode
    enc.load_state_dict(enc.state_dict())
    assert enc._uncond is None
    enc(["", "x"])
    assert enc._uncond is not None
    enc.train()
    enc(["", "x"])
    assert enc._uncond is None


def test_c_abi_library_exports_every_declared_symbol():
    """No compute calls here (no GPU): the library builds for sm_100a, loads, and exports what include/*.h declares."""
    path = B.build()
    lib = ctypes.CDLL(path)
    header = open(os.path.join(ROOT, "include", "ladiff_b200.h")).read()
    declared = set(re.findall(r"LADIFF_API\s+[\w\s\*]+?\b(ladiff_\w+)\s*\(", header))
    from ladiff_b200._lib import EXPORTS
    assert declared == set(EXPORTS), declared ^ set(EXPORTS)
    for s in declared:
        assert hasattr(lib, s), s
    lib.ladiff_abi_version.restype = ctypes.c_int
    assert lib.ladiff_abi_version() == 1
    if not torch.cuda.is_available():
        from ladiff_b200._lib import LadiffConfig
        h = ctypes.c_void_p(0)
        cfg = LadiffConfig(263, 9, 256, 4, 1024, 768, 5, 48, 196, 1)
        lib.ladiff_create.restype = ctypes.c_int
        assert lib.ladiff_create(ctypes.byref(cfg), ctypes.byref(h)) == -3      # LADIFF_ERR_CUDA: there is no CPU path
        lib.ladiff_last_error.restype = ctypes.c_char_p
        assert b"no CPU fallback" in lib.ladiff_last_error(None)


def test_bench_flop_model_matches_survey():
    import bench
    den, dec = bench.algorithmic_flops([196] * 128)
    assert abs((den + dec) / 1e12 - 2.224) < 0.002 and abs(den / 128 / 1e9 - 13.53) < 0.01 and abs(dec / 128 / 1e9 - 3.844) < 0.002
    lengths = O.synthetic_inputs(128, seed=1234, ragged=True)[2]
    assert sum(lengths) == 16092                                             # SURVEY.md 8d seeded ragged batch
    d2, e2 = bench.algorithmic_flops(lengths)
    assert abs((d2 + e2) / 1e12 - 1.376) < 0.002


def test_bench_algorithmic_flops_match_survey():
    """bench.py's roofline numerator is the SURVEY.md 8d contract figure: 27.03-27.07 MFLOP per valid latent row and step,
    13.53 + 3.84 = 17.38 GFLOP per prompt at 196 frames, 2.224 TFLOP per batch of 128; seeded ragged batch 1.376 TFLOP."""
    import importlib.util
    import numpy as np
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    den, dec = bench.algorithmic_flops([196])
    assert abs(den / 1e9 - 13.53) < 0.01 and abs(dec / 1e9 - 3.844) < 0.005
    assert abs((den + dec) * 128 / 1e12 - 2.224) < 0.002
    den1, _ = bench.algorithmic_flops([40], n_steps=1)           # m = 1: one valid row per CFG half
    assert abs(den1 / 2 / 1e6 - 27.03) < 0.01
    lengths = (np.random.default_rng(1234).integers(10, 50, size=128) * 4).tolist()
    assert sum(lengths) == 16092                                   # seeded ragged batch
    d, c = bench.algorithmic_flops(lengths)
    assert abs((d + c) / 1e12 - 1.376) < 0.002 and abs(c / 1e12 - 0.307) < 0.002


def test_sample_stream_pairing_rules():
    """Host logic of LADIFF.sample_stream's batch pairing (no GPU): which consecutive batches may share one reverse-loop call,
    and how their classifier-free-guidance halves are merged ([uncond | cond] per batch -> [uncond a, uncond b | cond a, cond b])."""
    from ladiff_b200.modeltype import LADIFF
    pairable = LADIFF._pairable.__get__(LADIFF.__new__(LADIFF))      # uses class constants only

    def item(n, cfg=True, lat=True, seed=0):
        g = torch.Generator().manual_seed(seed)
        t = torch.randn(((2 if cfg else 1) * n, 1, 768), generator=g)
        return t, [196] * n, (torch.randn((n, 5, 256), generator=g) if lat else None)

    assert pairable(item(128), item(128)) and pairable(item(80), item(80)) and pairable(item(177), item(177))
    assert not pairable(item(64), item(64))            # 128 prompts in one call: one chain -- nothing to gain
    assert not pairable(item(128), item(96))           # unequal batches would not map one batch to one chain
    assert not pairable(item(178), item(178))          # a chain of more than 177 prompts leaves the cluster feed-forward kernel
    assert not pairable(item(128), item(128, lat=False)) and not pairable(item(128), item(128, cfg=False))
    a, b = item(3, seed=1), item(3, seed=2)
    text, lengths, lat = LADIFF._merge_pair(a, b)
    assert lengths == a[1] + b[1] and torch.equal(lat, torch.cat([a[2], b[2]]))
    assert torch.equal(text[:3], a[0][:3]) and torch.equal(text[3:6], b[0][:3])            # unconditional halves first
    assert torch.equal(text[6:9], a[0][3:]) and torch.equal(text[9:], b[0][3:])
    text, _, lat = LADIFF._merge_pair(item(2, cfg=False, lat=False), item(2, cfg=False, lat=False, seed=5))
    assert text.shape[0] == 4 and lat is None
