#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SACB_VERBOSE=2 timeout -s KILL 200 python tools/sched_probe.py gen 128 6 6 1000 1 0.05 > gpurun_out/c17a.jsonl 2> gpurun_out/c17a.err
cat gpurun_out/c17a.jsonl; grep -c DDS gpurun_out/c17a.err; grep "DDS" gpurun_out/c17a.err | tail -2
