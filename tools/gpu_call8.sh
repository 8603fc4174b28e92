#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_grade.py -x -q 2>&1 | tail -15 > gpurun_out/c8_pytest.log
timeout 200 python tools/grade_probe.py 60000 128 > gpurun_out/c8_grade_probe.log 2>&1
cat gpurun_out/c8_pytest.log gpurun_out/c8_grade_probe.log
