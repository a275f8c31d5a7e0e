#!/bin/bash
# r01h: final round-1 numbers: full gpu suite, smoke, bench (own + reference arm), ncu --set full of the layer's kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s26_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s26_pytest.log
tail -3 gpurun_out/s26_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/s26_bench.json 2> gpurun_out/s26_bench.err
tail -c 1800 gpurun_out/s26_bench.json | head -c 900; echo; tail -2 gpurun_out/s26_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s26_ref.json 2>> gpurun_out/s26_bench.err
LADIFF_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_ffn_swap|k_attn_ln|k_linear_tc" -s 300 -c 3 \
   -o gpurun_out/s26_full python scripts/prof_step.py bf16x3 4 1 128 > gpurun_out/s26_ncufull.log 2>&1
tail -1 gpurun_out/s26_ncufull.log
