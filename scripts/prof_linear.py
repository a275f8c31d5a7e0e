"""Launches the fused linear on the hot-path shapes (for ncu captures and quick A/B timing)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ladiff_b200._lib import Engine, MODES

eng = Engine()
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["bf16x3", "bf16"]
shapes = {"den_qkv": (1280, 768, 256, "bias"), "den_out_ln": (1280, 256, 256, "ln"), "den_ffn1": (1280, 1024, 256, "gelu"),
          "den_ffn2_ln": (1280, 256, 1024, "ln"), "den_styl": (1280, 256, 1024, "ln_mod_silu"), "den_res": (1280, 256, 256, "res"),
          "dec_qkv": (25088, 768, 256, "bias"), "dec_ffn1": (25088, 1024, 256, "gelu"), "dec_ffn2_ln": (25088, 256, 1024, "ln"),
          "dec_out_ln": (25088, 256, 256, "ln")}
only = sys.argv[3].split(",") if len(sys.argv) > 3 else list(shapes)
for mode in modes:
    for name, (M, N, K, epi) in shapes.items():
        if name not in only:
            continue
        ms = eng.linear_bench(M, N, K, epi, MODES[mode], iters)
        print(f"{mode:7s} {name:12s} M={M:6d} N={N:5d} K={K:5d} {epi:12s} {ms*1e3:9.2f} us  {2.0*M*N*K/ms/1e9:8.1f} TFLOP/s")
