#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_grade.py -q 2>&1 | tail -40 > gpurun_out/c8_pytest.log
cat gpurun_out/c8_pytest.log
