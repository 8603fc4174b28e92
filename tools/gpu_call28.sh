#!/bin/bash
# final state of the round: smoke (now with the search-grade kernels), the search-grade tests after the dead-kernel removal, and the
# bench with NO flags (defaults: --steps 8 --warmup 3)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T0=$SECONDS
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c28_smoke.log 2>&1
echo "smoke wall $((SECONDS-T0)) s rc=$?"; tail -2 gpurun_out/c28_smoke.log | cut -c1-400
T0=$SECONDS
timeout -s KILL 200 python -m pytest tests/test_gpu_grade.py -m gpu -x -q > gpurun_out/c28_grade.log 2>&1
echo "grade tests wall $((SECONDS-T0)) s rc=$?"; tail -2 gpurun_out/c28_grade.log
T0=$SECONDS
timeout -s KILL 420 python bench.py > gpurun_out/bench_r2_n1_default.json 2> gpurun_out/c28_bench.err
echo "default bench wall $((SECONDS-T0)) s rc=$?"
wc -l gpurun_out/bench_r2_n1_default.json
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_r2_n1_default.json'))
    print({k:d[k] for k in ('value','ms_per_step','steps','bps','gpu_launches')}, d['cpu_baseline']['value'], d['bps_reference'])
except Exception as e: print("no json", e)
PY
