#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/s7_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s7_pytest.log
grep -E "max-abs|err|passed|failed|Error|exit" gpurun_out/s7_pytest.log | tail -30
for mode in bf16x3 bf16; do
  timeout 300 python scripts/prof_step.py $mode 50 5 128 >> gpurun_out/s7.log 2>&1
done
timeout 300 python scripts/prof_step.py bf16x3 50 3 1024 >> gpurun_out/s7.log 2>&1
cat gpurun_out/s7.log
