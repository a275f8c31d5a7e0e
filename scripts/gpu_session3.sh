#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s3_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s3_pytest.log
tail -15 gpurun_out/s3_pytest.log
timeout 300 python scripts/prof_linear.py 20 bf16x3,bf16 > gpurun_out/s3_linear.log 2>&1
cat gpurun_out/s3_linear.log
for mode in bf16x3 bf16; do
  LADIFF_CHAINS=1 timeout 300 python scripts/prof_step.py $mode 50 5 128 >> gpurun_out/s3_step.log 2>&1
done
LADIFF_CHAINS=1 LADIFF_ATTN_SIMT=1 timeout 300 python scripts/prof_step.py bf16x3 50 5 128 >> gpurun_out/s3_step.log 2>&1
LADIFF_CHAINS=1 timeout 300 python scripts/prof_step.py bf16x3 50 3 1024 >> gpurun_out/s3_step.log 2>&1
cat gpurun_out/s3_step.log
