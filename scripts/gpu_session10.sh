#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/s10_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s10_pytest.log
grep -E "max-abs|passed|failed|Error|exit" gpurun_out/s10_pytest.log | tail -16
for mode in bf16x3 bf16; do
  timeout 300 python scripts/prof_step.py $mode 50 5 128 >> gpurun_out/s10.log 2>&1
done
cat gpurun_out/s10.log
python scripts/trace_step.py bf16x3 50 128 2>&1 | tail -60 | head -24
