#!/bin/bash
# uniform-issue MMA/TMA warps in every tensor kernel + k_ffn_swap: full gpu suite, step timing, timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s18_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s18_pytest.log
tail -4 gpurun_out/s18_pytest.log
for env in "X=1" "LADIFF_NO_FFN_SWAP=1"; do
  echo "== $env" >> gpurun_out/s18.log
  env $env timeout 300 python scripts/prof_step.py bf16x3 50 5 128 >> gpurun_out/s18.log 2>&1
  env $env timeout 300 python scripts/prof_step.py bf16 50 5 128 >> gpurun_out/s18.log 2>&1
done
cat gpurun_out/s18.log
LADIFF_TRACE=1 timeout 300 python scripts/trace_step.py bf16x3 50 128 > gpurun_out/s18_trace.log 2>&1
head -12 gpurun_out/s18_trace.log; tail -2 gpurun_out/s18_trace.log
