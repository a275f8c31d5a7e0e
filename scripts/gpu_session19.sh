#!/bin/bash
# r01e: full gpu suite, bench (own + reference arm), ncu launch list of the bench command, ncu --set full of the loop's kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s19_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s19_pytest.log
tail -3 gpurun_out/s19_pytest.log
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/s19_bench.json 2> gpurun_out/s19_bench.err
tail -c 3300 gpurun_out/s19_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s19_ref.json 2>> gpurun_out/s19_bench.err
cat gpurun_out/s19_ref.json
LADIFF_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s19_launches.csv python bench.py --steps 1 --warmup 1 --quick > gpurun_out/s19_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/s19_launches.csv | head -16
LADIFF_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_ffn_swap|k_attn_ln|k_linear_tc" -s 300 -c 6 \
   -o gpurun_out/s19_full python scripts/prof_step.py bf16x3 4 1 128 > gpurun_out/s19_ncufull.log 2>&1
tail -2 gpurun_out/s19_ncufull.log
