#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s6_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s6_pytest.log
tail -5 gpurun_out/s6_pytest.log
for env in "X=1" "LADIFF_NO_PDL=1"; do
  for mode in bf16x3 bf16; do
  echo "== $env $mode" >> gpurun_out/s6.log
  env $env timeout 300 python scripts/prof_step.py $mode 50 5 128 >> gpurun_out/s6.log 2>&1
  done
done
timeout 300 python scripts/prof_step.py bf16x3 50 3 1024 >> gpurun_out/s6.log 2>&1
cat gpurun_out/s6.log
timeout 300 python scripts/prof_linear.py 20 bf16x3 den_qkv,den_out_ln,den_ffn1,den_ffn2_ln,den_styl,den_res > gpurun_out/s6_linear.log 2>&1
cat gpurun_out/s6_linear.log
