#!/bin/bash
set -x
mkdir -p gpurun_out
LADIFF_DBG_STAMPS=1 timeout 300 python scripts/prof_linear.py 20 bf16x3 > gpurun_out/s4_stamps.log 2>&1
cat gpurun_out/s4_stamps.log
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:k_linear_tc -s 5 -c 1 -o gpurun_out/prof2_dec_ffn1 -f python scripts/prof_linear.py 4 bf16x3 dec_ffn1 > gpurun_out/s4_b.log 2>&1
timeout 600 $NCU -k regex:k_linear_tc -s 5 -c 1 -o gpurun_out/prof2_dec_qkv -f python scripts/prof_linear.py 4 bf16x3 dec_qkv > gpurun_out/s4_c.log 2>&1
