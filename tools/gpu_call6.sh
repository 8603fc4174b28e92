#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SACB_SG_PARTS=3 timeout 100 python tools/sg_debug.py 40000 2>&1 | grep "k=" > gpurun_out/c6_sg_debug.log
timeout 300 python -m pytest tests/test_gpu_grade.py -x -q 2>&1 | tail -15 > gpurun_out/c6_pytest.log
timeout 200 python tools/grade_probe.py 60000 128 > gpurun_out/c6_grade_probe.log 2>&1
cat gpurun_out/c6_sg_debug.log gpurun_out/c6_pytest.log gpurun_out/c6_grade_probe.log
