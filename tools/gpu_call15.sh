#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 200 python -m pytest tests/test_gpu_grade.py -x -q 2>&1 | tail -6 | cut -c1-500 > gpurun_out/c15_pytest.log
cat gpurun_out/c15_pytest.log
if grep -q "6 passed" gpurun_out/c15_pytest.log; then
  timeout -s KILL 700 python bench.py --steps 10 --warmup 3 --gen 128 --inflight 10 > gpurun_out/c15_bench.json 2> gpurun_out/c15_bench.err
  python - <<'PY'
import json
d=json.load(open('gpurun_out/c15_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','bps','gpu_launches')}, d['chains'], d['device_ms_by_kernel_class'], d['kernel_ms_per_generation'])
PY
  tail -3 gpurun_out/c15_bench.err
fi
