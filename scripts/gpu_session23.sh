#!/bin/bash
mkdir -p gpurun_out
LADIFF_TRACE=1 timeout 300 python scripts/trace_step.py bf16x3 50 128 > gpurun_out/s23_trace.log 2>&1
head -12 gpurun_out/s23_trace.log; tail -2 gpurun_out/s23_trace.log
