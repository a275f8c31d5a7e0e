"""The cluster-fused feed-forward kernel (csrc/ffn_swap.cuh: token groups of <= 48 on the MMA N axis, clusters of 4)
against a torch fp64 restatement of the two feed-forward
pairs of a denoiser layer (reference: mdiff_transformer.py:60-62 sa_block FFN + norm2; :137-162,248-262 FFN + StylizationBlock
prologue), and against the four separate fused linears it replaces."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BLOCKS = ["input_blocks.0", "input_blocks.1", "input_blocks.2", "input_blocks.3", "middle_block",
          "output_blocks.0", "output_blocks.1", "output_blocks.2", "output_blocks.3"]
TOL = {"bf16x3": 1e-4, "bf16": 4e-2}


def ref_ffn(sd, layer, x, mod):
    P = f"denoiser.encoder.{BLOCKS[layer]}."
    w = lambda k: sd[P + k].double()
    x = x.double()
    h = F.relu(F.linear(x, w("sa_block.linear1.weight"), w("sa_block.linear1.bias")))
    y = x + F.linear(h, w("sa_block.linear2.weight"), w("sa_block.linear2.bias"))
    x3 = F.layer_norm(y, (256,), w("sa_block.norm2.weight"), w("sa_block.norm2.bias"), 1e-5)
    h = F.gelu(F.linear(x3, w("ffn.linear1.weight"), w("ffn.linear1.bias")))
    y = F.linear(h, w("ffn.linear2.weight"), w("ffn.linear2.bias"))
    y = F.layer_norm(y, (256,), w("ffn.proj_out.norm.weight"), w("ffn.proj_out.norm.bias"), 1e-5)
    s = F.silu(y * (1 + mod[:256].double()) + mod[256:].double())
    return x3, s


@pytest.mark.parametrize("mode", ["bf16x3", "bf16"])
@pytest.mark.parametrize("M,layer", [(1280, 0), (128, 4), (77, 8), (1, 2), (300, 5), (2000, 7)])
def test_ffn_fused(engine, oracle_sd, mode, M, layer):
    from ladiff_b200._lib import MODES
    g = torch.Generator(device="cpu").manual_seed(M * 31 + layer)
    x = torch.randn((M, 256), generator=g)
    mod = 0.3 * torch.randn((512,), generator=g)
    x3r, sr = ref_ffn(oracle_sd, layer, x, mod)
    x3, s, _ = engine.ffn_test(x.cuda(), layer, mod.cuda(), mode=MODES[mode], fused=True)    # k_ffn_swap for M <= 1776
    x3u, su, _ = engine.ffn_test(x.cuda(), layer, mod.cuda(), mode=MODES[mode], fused=False)
    assert torch.isfinite(x3).all() and torch.isfinite(s).all()
    for name, got, ref in (("x3", x3, x3r), ("s", s, sr), ("x3 unfused", x3u, x3r), ("s unfused", su, sr)):
        err = (got.double().cpu() - ref).abs().max().item() / ref.abs().max().item()
        assert err < TOL[mode], f"{mode} M={M} layer={layer} {name}: rel max err {err:.3e}"


@pytest.mark.parametrize("mode", ["bf16x3", "bf16"])
@pytest.mark.parametrize("M,layer", [(1777, 1), (4173, 6), (25088, 3)])
def test_ffn_tile(engine, oracle_sd, mode, M, layer):
    """Above 1776 rows every feed-forward pair runs as ONE persistent tile kernel (csrc/ffn_tile.cuh: hidden activation kept in
    tensor memory as the A operand of the second GEMM): both epilogue kinds (ReLU + LN + residual; GELU + LN * mod + SiLU),
    a ragged last tile, one and two tiles per CTA (25 088 rows = 196 tiles on 148 SMs)."""
    from ladiff_b200._lib import MODES
    g = torch.Generator(device="cpu").manual_seed(M * 7 + layer)
    x = torch.randn((M, 256), generator=g)
    mod = 0.3 * torch.randn((512,), generator=g)
    x3r, sr = ref_ffn(oracle_sd, layer, x, mod)
    x3, s, _ = engine.ffn_test(x.cuda(), layer, mod.cuda(), mode=MODES[mode], fused=True)
    assert torch.isfinite(x3).all() and torch.isfinite(s).all()
    for name, got, ref in (("x3", x3, x3r), ("s", s, sr)):
        err = (got.double().cpu() - ref).abs().max().item() / ref.abs().max().item()
        assert err < TOL[mode], f"{mode} M={M} layer={layer} {name}: rel max err {err:.3e}"
    x3b, sb, _ = engine.ffn_test(x.cuda(), layer, mod.cuda(), mode=MODES[mode], fused=True)
    assert torch.equal(x3, x3b) and torch.equal(s, sb)


def test_ffn_fused_deterministic(engine):
    from ladiff_b200._lib import MODES
    g = torch.Generator(device="cpu").manual_seed(5)
    x = torch.randn((1280, 256), generator=g).cuda()
    mod = (0.3 * torch.randn((512,), generator=g)).cuda()
    a = engine.ffn_test(x, 3, mod, mode=MODES["bf16x3"], fused=True)
    b = engine.ffn_test(x, 3, mod, mode=MODES["bf16x3"], fused=True)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])   # fixed-order reduce-scatter: bit-reproducible


@pytest.mark.parametrize("rt", [16, 32, 48])
def test_ffn_swap_group_sizes(engine, oracle_sd, rt, monkeypatch):
    """every token-group size of k_ffn_swap (the launcher picks it from the row count; forced here)"""
    from ladiff_b200._lib import MODES
    monkeypatch.setenv("LADIFF_FFN_RT", str(rt))
    g = torch.Generator(device="cpu").manual_seed(rt)
    for M in (rt * 3 + 5, 1):
        x = torch.randn((M, 256), generator=g)
        mod = 0.3 * torch.randn((512,), generator=g)
        x3r, sr = ref_ffn(oracle_sd, 6, x, mod)
        x3, s, _ = engine.ffn_test(x.cuda(), 6, mod.cuda(), mode=MODES["bf16x3"], fused=True)
        for name, got, ref in (("x3", x3, x3r), ("s", s, sr)):
            err = (got.double().cpu() - ref).abs().max().item() / ref.abs().max().item()
            assert err < TOL["bf16x3"], f"rt={rt} M={M} {name}: rel max err {err:.3e}"
