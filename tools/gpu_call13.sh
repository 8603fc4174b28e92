#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SACB_VERBOSE=2 timeout -s KILL 140 python tools/sched_probe.py gen 128 6 6 1000 1 0.05 > gpurun_out/c13a.jsonl 2> gpurun_out/c13a.err
SACB_VERBOSE=2 timeout -s KILL 140 python tools/sched_probe.py gen 128 6 6 1000 0 0.05 > gpurun_out/c13b.jsonl 2> gpurun_out/c13b.err
cat gpurun_out/c13a.jsonl gpurun_out/c13b.jsonl; for f in a b; do grep -c DDS gpurun_out/c13$f.err; grep "DDS" gpurun_out/c13$f.err | tail -3; done
