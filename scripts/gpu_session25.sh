#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ffn.py tests/test_gpu_linear.py -m gpu -x -q > gpurun_out/s25_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s25_pytest.log
tail -3 gpurun_out/s25_pytest.log
LADIFF_TRACE=1 timeout 300 python scripts/trace_step.py bf16 50 128 > gpurun_out/s25_trace_bf16.log 2>&1
head -9 gpurun_out/s25_trace_bf16.log; tail -2 gpurun_out/s25_trace_bf16.log
timeout 300 python scripts/prof_step.py bf16x3 50 3 1024 2>&1 | tail -2
