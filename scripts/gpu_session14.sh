#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/s14_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s14_pytest.log
grep -E "max-abs|passed|failed|Error|exit" gpurun_out/s14_pytest.log | tail -16
for env in "X=1" "LADIFF_NO_FFN_CLUSTER=1"; do
  echo "== $env" >> gpurun_out/s14.log
  env $env timeout 300 python scripts/prof_step.py bf16x3 50 5 128 >> gpurun_out/s14.log 2>&1
  env $env timeout 300 python scripts/prof_step.py bf16 50 5 128 >> gpurun_out/s14.log 2>&1
done
cat gpurun_out/s14.log
LADIFF_TRACE=1 python scripts/trace_step.py bf16x3 50 128 2>&1 | tail -40 | head -14
