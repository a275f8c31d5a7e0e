"""One fused-linear shape for an ncu capture: python scripts/ncu_linear.py M N K epilogue [mode]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ladiff_b200._lib import Engine, MODES
M, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
epi = sys.argv[4]
mode = sys.argv[5] if len(sys.argv) > 5 else "bf16x3"
eng = Engine(nfeats=263)
ms = eng.linear_bench(M, N, K, epi, MODES[mode], 3)
print(f"M={M} N={N} K={K} {epi} {mode}: {ms * 1e3:.2f} us per launch")
