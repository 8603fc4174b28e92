#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 560 python bench.py --steps 6 --warmup 3 --gen 128 --inflight 6 > gpurun_out/c14_bench.json 2> gpurun_out/c14_bench.err
tail -c 3500 gpurun_out/c14_bench.json; tail -5 gpurun_out/c14_bench.err
