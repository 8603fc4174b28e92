#!/bin/bash
# config C5 at the real window on one GPU: one heterogeneous DDS generation (sigma 0.25 around the default) of 64 / 256 / 1024 candidates
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 55 python tools/sweep_probe.py --pops 64,512 --window 441000 --grade 1 > gpurun_out/sweep_r2_n1.jsonl 2> gpurun_out/c29.err
echo "rc=$?"; cat gpurun_out/sweep_r2_n1.jsonl; tail -2 gpurun_out/c29.err | cut -c1-300
