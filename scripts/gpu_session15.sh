#!/bin/bash
# state check after re-entry: gpu tests, FFN-cluster A/B, per-launch timeline, bench, ncu --set full of the loop's top kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s15_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s15_pytest.log
tail -4 gpurun_out/s15_pytest.log
for env in "X=1" "LADIFF_NO_FFN_CLUSTER=1"; do
  echo "== $env" >> gpurun_out/s15.log
  env $env timeout 300 python scripts/prof_step.py bf16x3 50 5 128 >> gpurun_out/s15.log 2>&1
  env $env timeout 300 python scripts/prof_step.py bf16 50 5 128 >> gpurun_out/s15.log 2>&1
done
cat gpurun_out/s15.log
LADIFF_TRACE=1 timeout 300 python scripts/trace_step.py bf16x3 50 128 > gpurun_out/s15_trace.log 2>&1
head -40 gpurun_out/s15_trace.log
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/s15_bench.json 2> gpurun_out/s15_bench.err
tail -c 3500 gpurun_out/s15_bench.json
LADIFF_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_ffn_cluster|k_attn_ln|k_linear_tc" -s 300 -c 6 \
   -o gpurun_out/s15_full python scripts/prof_step.py bf16x3 4 1 128 > gpurun_out/s15_ncu.log 2>&1
tail -3 gpurun_out/s15_ncu.log
ls -la gpurun_out
