#!/bin/bash
mkdir -p gpurun_out
env | grep -i nccl | head -5
for i in 1 2; do
for e in "X=1" "LADIFF_LN_CLUSTER_ALL=1"; do
  echo "== $e"; env $e timeout 300 python scripts/prof_step.py bf16x3 50 10 128 2>&1 | tail -1
done
done
