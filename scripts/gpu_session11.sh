#!/bin/bash
# round-1 refresh of the committed numbers: bench (own arm + reference arm), ncu launch list of the bench command
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/s11_bench.json 2> gpurun_out/s11_bench.err
tail -c 3000 gpurun_out/s11_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s11_ref.json 2>> gpurun_out/s11_bench.err
cat gpurun_out/s11_ref.json
LADIFF_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s11_launches.csv python bench.py --steps 1 --warmup 1 --quick > gpurun_out/s11_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/s11_launches.csv | head -30
