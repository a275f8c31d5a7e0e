#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s24_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s24_pytest.log
tail -3 gpurun_out/s24_pytest.log
timeout 300 python scripts/prof_step.py bf16x3 50 5 128 2>&1 | tail -2
timeout 300 python scripts/prof_step.py bf16 50 5 128 2>&1 | tail -2
LADIFF_TRACE=1 timeout 300 python scripts/trace_step.py bf16x3 50 128 > gpurun_out/s24_trace.log 2>&1
head -9 gpurun_out/s24_trace.log; tail -2 gpurun_out/s24_trace.log
