#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SACB_SG_PARTS=1 timeout -s KILL 60 python tools/sg_debug.py 20000 2>&1 | grep "k=" > gpurun_out/c18_dbg.log
timeout -s KILL 150 python -m pytest tests/test_gpu_grade.py -x -q 2>&1 | tail -6 | cut -c1-600 > gpurun_out/c18_pytest.log
timeout -s KILL 100 python tools/grade_probe.py 60000 128 > gpurun_out/c18_grade_probe.log 2>&1
cat gpurun_out/c18_dbg.log gpurun_out/c18_pytest.log gpurun_out/c18_grade_probe.log
