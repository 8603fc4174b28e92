#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T0=$SECONDS
SACB_TRACE_DDS=/tmp/bench_trace.jsonl timeout -s KILL 870 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err
echo "bench wall $((SECONDS-T0)) s rc=$?"
head -1100 /tmp/bench_trace.jsonl > gpurun_out/bench_trace_first.jsonl 2>/dev/null
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_r2_n1.json'))
    print({k:d[k] for k in ('value','ms_per_step','bps','gpu_launches')}, d['chains'], d['device_ms_by_kernel_class'], d['kernel_ms_per_generation'], d['cpu_baseline']['value'], d['clocks'])
except Exception as e: print("no json", e)
PY
tail -3 gpurun_out/bench_r2_n1.err | cut -c1-300
