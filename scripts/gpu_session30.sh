#!/bin/bash
mkdir -p gpurun_out
for shape in "25088 256 256 ln" "25088 1024 256 gelu" "25088 768 256 bias" "25088 256 1024 ln"; do python scripts/ncu_linear.py $shape 2>&1 | tail -1; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_linear_tc -s 4 -c 1 -o gpurun_out/s30_ln python scripts/ncu_linear.py 25088 256 256 ln > gpurun_out/s30_ncu.log 2>&1; tail -1 gpurun_out/s30_ncu.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_linear_tc -s 4 -c 1 -o gpurun_out/s30_gelu python scripts/ncu_linear.py 25088 1024 256 gelu > gpurun_out/s30_ncu2.log 2>&1; tail -1 gpurun_out/s30_ncu2.log
