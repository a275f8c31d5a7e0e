#!/bin/bash
# ncu --set full captures of the fused linear at two shapes + the decoder self-attention
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:k_linear_tc -s 5 -c 1 -o gpurun_out/prof_den_qkv -f python scripts/prof_linear.py 4 bf16x3 den_qkv > gpurun_out/s2_a.log 2>&1
timeout 600 $NCU -k regex:k_linear_tc -s 5 -c 1 -o gpurun_out/prof_dec_ffn1 -f python scripts/prof_linear.py 4 bf16x3 dec_ffn1 > gpurun_out/s2_b.log 2>&1
timeout 600 $NCU -k regex:k_linear_tc_ln -s 5 -c 1 -o gpurun_out/prof_den_ffn2_ln -f python scripts/prof_linear.py 4 bf16x3 den_ffn2_ln > gpurun_out/s2_c.log 2>&1
LADIFF_NO_GRAPH=1 timeout 900 $NCU -k regex:k_attn_self -s 2 -c 1 -o gpurun_out/prof_attn_self -f python scripts/prof_step.py bf16x3 2 1 128 > gpurun_out/s2_d.log 2>&1
tail -3 gpurun_out/s2_*.log
ls -la gpurun_out
