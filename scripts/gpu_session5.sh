#!/bin/bash
mkdir -p gpurun_out
for env in "X=1" "LADIFF_NO_PDL=1" "LADIFF_NO_GRAPH=1" "LADIFF_NO_CLUSTER=1" "LADIFF_TILE_SMS=296" "LADIFF_TILE_SMS=74"; do
  echo "== $env" >> gpurun_out/s5.log
  env $env LADIFF_CHAINS=1 timeout 300 python scripts/prof_step.py bf16x3 50 5 128 >> gpurun_out/s5.log 2>&1
done
cat gpurun_out/s5.log
