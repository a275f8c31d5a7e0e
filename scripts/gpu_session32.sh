#!/bin/bash
# r01i: final round-1 evidence: gpu suite, smoke, bench (own + reference arm), ncu launch list of the bench command
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "passed|failed|error" | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/s32_bench.json 2> gpurun_out/s32_bench.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/s32_bench.json') if l.startswith('{')][-1])
print({k: d[k] for k in ('value','ms_per_step','latency_ms_per_batch','value_sequential','reverse_ms','decode_ms','value_bf16')}); print(d['e2e']); print(d['ragged']); print({k: d['roofline'][k] for k in ('achieved','frac','us_per_launch','share_of_step')})"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s32_ref.json 2>> gpurun_out/s32_bench.err
head -c 300 gpurun_out/s32_ref.json; echo
LADIFF_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s32_launches.csv python bench.py --steps 1 --warmup 1 --quick --no-pipeline > gpurun_out/s32_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/s32_launches.csv 2>/dev/null | head -12
