#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T0=$SECONDS
timeout -s KILL 860 python bench.py --steps 20 --warmup 5 --gen 128 --inflight 20 > gpurun_out/c20_bench.json 2> gpurun_out/c20_bench.err
echo "bench wall $((SECONDS-T0)) s rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/c20_bench.json'))
    print({k:d[k] for k in ('value','ms_per_step','bps','gpu_launches')}, d['chains'], d['device_ms_by_kernel_class'], d['kernel_ms_per_generation'], d['cpu_baseline']['value'])
except Exception as e: print("no json", e)
PY
tail -3 gpurun_out/c20_bench.err | cut -c1-300
