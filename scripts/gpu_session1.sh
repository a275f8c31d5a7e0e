#!/bin/bash
# GPU session: parity tests, chain sweep, stamps
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s1_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s1_pytest.log
tail -5 gpurun_out/s1_pytest.log
for ch in 1 2 4 8; do
  for mode in bf16x3 bf16; do
    echo "== chains $ch mode $mode" >> gpurun_out/s1_chains.log
    LADIFF_CHAINS=$ch timeout 300 python scripts/prof_step.py $mode 50 5 128 >> gpurun_out/s1_chains.log 2>&1
  done
done
LADIFF_CHAINS=4 timeout 300 python scripts/prof_step.py bf16x3 50 3 1024 >> gpurun_out/s1_chains.log 2>&1
LADIFF_CHAINS=1 timeout 300 python scripts/prof_step.py bf16x3 50 3 1024 >> gpurun_out/s1_chains.log 2>&1
cat gpurun_out/s1_chains.log
LADIFF_DBG_STAMPS=1 timeout 300 python scripts/prof_linear.py 20 bf16x3 > gpurun_out/s1_stamps.log 2>&1
cat gpurun_out/s1_stamps.log
