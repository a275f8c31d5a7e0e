#!/bin/bash
# sample_stream (pipelined) test + bench with the dominant-kernel roofline
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ffn.py -m gpu -x -q > gpurun_out/s21_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/s21_pytest.log
tail -4 gpurun_out/s21_pytest.log
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/s21_bench.json 2> gpurun_out/s21_bench.err
tail -c 4500 gpurun_out/s21_bench.json; tail -3 gpurun_out/s21_bench.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
