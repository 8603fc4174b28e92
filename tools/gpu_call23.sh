#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 900 python -m pytest tests -m gpu -x -q --durations=12 2>&1 | tail -30 | cut -c1-300 > gpurun_out/c23_pytest.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 >> gpurun_out/c23_pytest.log
cat gpurun_out/c23_pytest.log
