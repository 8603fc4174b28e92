#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SACB_VERBOSE=2 timeout -s KILL 150 python tools/sched_probe.py spec 16 4 4 1000 1 0.05 > gpurun_out/c12_sched.jsonl 2> gpurun_out/c12_sched.err
cat gpurun_out/c12_sched.jsonl; grep "frame 0" gpurun_out/c12_sched.err | awk 'NR%10==1' | tail -30; grep -c "frame 0" gpurun_out/c12_sched.err
