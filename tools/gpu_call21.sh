#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 400 python -m pytest tests/test_gpu_search_bytes.py -x -q -s 2>&1 | tail -12 | cut -c1-400 > gpurun_out/c21_bytes.log
timeout -s KILL 200 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -8 | cut -c1-400 > gpurun_out/c21_multi.log
cat gpurun_out/c21_bytes.log gpurun_out/c21_multi.log
