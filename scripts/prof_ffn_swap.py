"""k_ffn_swap in isolation: back-to-back timing and (LADIFF_DBG_STAMPS=1) per-CTA clock stamps of one launch.
    LADIFF_DBG_STAMPS=1 python scripts/prof_ffn_swap.py [M=1280] [mode=bf16x3]
Synthetic weights (tests' generator; the oracle is only used here as the weight generator of the test fixtures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ladiff_b200._lib import Engine, MODES
from oracle import ladiff_oracle as O

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1280
mode = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
sd = O.make_state_dict(1234, 263, perturb=True)
eng = Engine(nfeats=263)
eng.set_weights({k: v.cuda() for k, v in O.sub(sd, "denoiser.").items()}, "denoiser.")
eng.set_weights({k: v.cuda() for k, v in O.sub(sd, "vae.").items()}, "vae.")
eng.finalize(3)
g = torch.Generator().manual_seed(0)
x = torch.randn((M, 256), generator=g).cuda()
mod = (0.1 * torch.randn((512,), generator=g)).cuda()
for _ in range(2):
    _, _, ms = eng.ffn_test(x, 3, mod, mode=MODES[mode], fused=True, iters=200)
print(f"M={M} {mode} fused ffn: {ms * 1e3:.2f} us per call (back-to-back)")
