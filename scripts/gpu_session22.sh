#!/bin/bash
# fused attention prologue in k_ffn_swap: parity, step timing A/B, timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s22_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s22_pytest.log
tail -6 gpurun_out/s22_pytest.log
for env in "X=1" "LADIFF_NO_ATT_FUSE=1"; do
  echo "== $env" >> gpurun_out/s22.log
  env $env timeout 300 python scripts/prof_step.py bf16x3 50 5 128 >> gpurun_out/s22.log 2>&1
  env $env timeout 300 python scripts/prof_step.py bf16 50 5 128 >> gpurun_out/s22.log 2>&1
done
cat gpurun_out/s22.log
LADIFF_TRACE=1 timeout 300 python scripts/trace_step.py bf16x3 50 128 > gpurun_out/s22_trace.log 2>&1
head -9 gpurun_out/s22_trace.log; tail -2 gpurun_out/s22_trace.log
