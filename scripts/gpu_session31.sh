#!/bin/bash
# robustness: the gpu suite three times, memcheck on the cluster-kernel tests; decoder LN GEMM timing after the residual prefetch
mkdir -p gpurun_out
for i in 1 2 3; do timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -1; done
for shape in "25088 256 256 ln" "25088 256 1024 ln"; do python scripts/ncu_linear.py $shape 2>&1 | tail -1; done
timeout 300 python scripts/prof_step.py bf16x3 50 10 128 2>&1 | tail -2
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_ffn.py -m gpu -x -q -k "swap_group_sizes or deterministic" > gpurun_out/s31_memcheck.log 2>&1; tail -5 gpurun_out/s31_memcheck.log
