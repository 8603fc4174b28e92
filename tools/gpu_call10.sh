#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 250 python -m pytest tests/test_gpu_grade.py -x -q 2>&1 | tail -8 | cut -c1-400 > gpurun_out/c10_pytest.log
{ timeout -s KILL 330 python tools/sched_probe.py spec 16 8 8 100 1
  timeout -s KILL 330 python tools/sched_probe.py spec 16 8 8 100 0
} > gpurun_out/c10_sched.jsonl 2> gpurun_out/c10_sched.err
cat gpurun_out/c10_pytest.log gpurun_out/c10_sched.jsonl; tail -3 gpurun_out/c10_sched.err
