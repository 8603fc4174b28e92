#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 400 python -m pytest tests/test_gpu_grade.py tests/test_gpu_multi.py tests/test_gpu_parity.py tests/test_gpu_sparse.py -m gpu -x -q 2>&1 | tail -15 | cut -c1-300 > gpurun_out/c24_pytest.log
cat gpurun_out/c24_pytest.log
