#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SACB_VERBOSE=0 timeout -s KILL 300 python -m pytest tests/test_gpu_search_bytes.py -x -q -s 2>&1 | tail -12 | cut -c1-500 > gpurun_out/c22_bytes.log
cat gpurun_out/c22_bytes.log
