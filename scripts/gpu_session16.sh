#!/bin/bash
# k_ffn_swap bring-up: parity tests, per-call timing + stamps, then the step A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ffn.py -m gpu -x -q > gpurun_out/s16_ffn.log 2>&1; echo "ffn pytest exit $?" >> gpurun_out/s16_ffn.log
tail -25 gpurun_out/s16_ffn.log
LADIFF_DBG_STAMPS=1 timeout 120 python - <<'PY' 2>&1 | tee gpurun_out/s16_time.log
import torch, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from ladiff_b200._lib import Engine, MODES
from oracle import ladiff_oracle as O
sd = O.make_state_dict(1234, 263, perturb=True)
eng = Engine(nfeats=263)
eng.set_weights({k: v.cuda() for k, v in O.sub(sd, "denoiser.").items()}, "denoiser.")
eng.finalize(1)
x = torch.randn(1280, 256).cuda(); mod = (0.3 * torch.randn(512)).cuda()
for mode in ("bf16x3", "bf16"):
    for fused in (1, 2):
        _, _, ms = eng.ffn_test(x, 3, mod, mode=MODES[mode], fused=fused, iters=200)
        print(f"{mode} fused={fused}: {ms*1e3:.2f} us per call (M=1280, back-to-back)")
PY
for env in "X=1" "LADIFF_NO_FFN_SWAP=1"; do
  echo "== $env" >> gpurun_out/s16.log
  env $env timeout 300 python scripts/prof_step.py bf16x3 50 5 128 >> gpurun_out/s16.log 2>&1
  env $env timeout 300 python scripts/prof_step.py bf16 50 5 128 >> gpurun_out/s16.log 2>&1
done
cat gpurun_out/s16.log
