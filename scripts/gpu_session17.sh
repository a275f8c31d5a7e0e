#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ffn.py -m gpu -x -q > gpurun_out/s17_ffn.log 2>&1; echo "ffn pytest exit $?" >> gpurun_out/s17_ffn.log
tail -5 gpurun_out/s17_ffn.log
LADIFF_DBG_STAMPS=1 timeout 120 python - <<'PY' 2>&1 | tee gpurun_out/s17_time.log
import torch, sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from ladiff_b200._lib import Engine, MODES
from oracle import ladiff_oracle as O
sd = O.make_state_dict(1234, 263, perturb=True)
eng = Engine(nfeats=263)
eng.set_weights({k: v.cuda() for k, v in O.sub(sd, "denoiser.").items()}, "denoiser.")
eng.finalize(1)
mod = (0.3 * torch.randn(512)).cuda()
for M in (1280, 48, 192):
    x = torch.randn(M, 256).cuda()
    for mode in ("bf16x3", "bf16"):
        os.environ["LADIFF_FFN_RT"] = "48"
        _, _, ms = eng.ffn_test(x, 3, mod, mode=MODES[mode], fused=1, iters=200)
        print(f"M={M} {mode} swap: {ms*1e3:.2f} us per call (back-to-back)")
PY
