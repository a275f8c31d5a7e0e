#!/bin/bash
# A/B on one box: builds selected with LADIFF_LIB
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ffn.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | grep -E "passed|failed|Error" | tail -3
for i in 1 2; do
for lib in "" prev; do
  p=""; [ -n "$lib" ] && p="ladiff_b200/_C/libladiff_b200_$lib.so"
  echo "== LADIFF_LIB=$p" | tee -a gpurun_out/s27.log
  LADIFF_LIB=$p timeout 300 python scripts/prof_step.py bf16x3 50 10 128 2>&1 | tail -2 | head -1 | tee -a gpurun_out/s27.log
done
done
