#!/bin/bash
# One GPU-box session: the gpu test suite, smoke, bench (own + reference arm), optional traces.  Usage (from the repo root):
#   gpurun --timeout 2400 -- 'bash scripts/gpu_round.sh <tag> [tests|bench|trace|ncu ...]'
# Everything lands under gpurun_out/<tag>_*; copy what should be judged into profiles/.
tag=${1:-s}; shift
what=${*:-tests bench}
mkdir -p gpurun_out
for w in $what; do
  case $w in
    tests)
      timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/${tag}_pytest.log 2>&1
      grep -E "passed|failed|error|measured|max-abs err" gpurun_out/${tag}_pytest.log | tail -60
      python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 ;;
    bench)
      timeout 1200 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
      tail -c 3000 gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err ;;
    quick)   # A/B timing: reverse / decode split + the pipelined bench line without extras
      python scripts/prof_step.py bf16x3 50 5 128 2>&1 | tail -2
      timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --quick > gpurun_out/${tag}_quick.json 2>> gpurun_out/${tag}_bench.err
      python -c "
import json; d=json.loads([l for l in open('gpurun_out/${tag}_quick.json') if l.startswith('{')][-1])
print({k: d.get(k) for k in ('value','ms_per_step','latency_ms_per_batch','value_sequential','p50_latency_ms')}, d['e2e']['value'])" ;;
    ref)
      timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_ref.json 2>> gpurun_out/${tag}_bench.err
      head -c 600 gpurun_out/${tag}_ref.json; echo ;;
    trace)
      for B in 8 32 128; do
        timeout 300 python scripts/trace_step.py bf16x3 50 $B > gpurun_out/${tag}_trace_B$B.txt 2>&1
        head -14 gpurun_out/${tag}_trace_B$B.txt | tail -9
      done ;;
    variants)   # opt-in / alternate kernel instantiations, selected by environment switches read when a plan is captured
      for v in LADIFF_ATTN_LN_MAXT8=1 LADIFF_ATTN_2SEQ=1 LADIFF_ATTN_LN_128=1 LADIFF_ATTN_F32_STAGE=1 LADIFF_NO_FFN_TILE=1 LADIFF_ATT_FUSE=1; do
        echo "== $v"; env $v timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sampling_vs_reference or batch32 or decode_vs_reference or encode" 2>&1 | tail -1
      done ;;
    launches)
      LADIFF_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv \
        python bench.py --steps 1 --warmup 1 --quick --no-pipeline > gpurun_out/${tag}_ncu.log 2>&1
      python scripts/summarize_launches.py gpurun_out/${tag}_launches.csv 2>/dev/null | head -24 ;;
  esac
done
