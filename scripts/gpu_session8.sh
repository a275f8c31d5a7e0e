#!/bin/bash
mkdir -p gpurun_out
LADIFF_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/s8_launches_warm.csv python scripts/prof_step.py bf16x3 3 1 128 > gpurun_out/s8.log 2>&1
python scripts/summarize_launches.py gpurun_out/s8_launches_warm.csv
