#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
for e in "X=1" "LADIFF_ATTN_2SEQ=1"; do
  echo "== $e"; env $e timeout 300 python scripts/prof_step.py bf16x3 50 10 128 2>&1 | tail -2 | head -1
done
done
LADIFF_ATTN_2SEQ=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
