#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{ timeout -s KILL 200 python tools/sched_probe.py spec 16 8 8 1000 1 0.05
  timeout -s KILL 200 python tools/sched_probe.py spec 32 8 8 1000 1 0.05
  timeout -s KILL 200 python tools/sched_probe.py gen 128 8 8 1000 1 0.05
  timeout -s KILL 200 python tools/sched_probe.py spec 16 8 8 1000 0 0.05
} > gpurun_out/c11_sched.jsonl 2> gpurun_out/c11_sched.err
cat gpurun_out/c11_sched.jsonl; tail -3 gpurun_out/c11_sched.err
