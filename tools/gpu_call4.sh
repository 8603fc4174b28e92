#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SACB_SG_PARTS=3 timeout 100 python tools/sg_debug.py 40000 2>&1 | grep "k=4" > gpurun_out/c4_sg_debug.log
timeout 300 python -m pytest tests/test_gpu_grade.py -x -q 2>&1 | tail -15 > gpurun_out/c4_pytest.log
timeout 200 python tools/grade_probe.py 60000 128 > gpurun_out/c4_grade_probe.log 2>&1
SACB_GRADE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:cascade_sg -c 1 -o gpurun_out/casc_sg_r2a -f python tools/prof_run.py 8 6000 > gpurun_out/c4_ncu1.log 2>&1
SACB_GRADE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:ols_sg -c 1 -o gpurun_out/ols_sg_r2a -f python tools/prof_run.py 8 6000 > gpurun_out/c4_ncu2.log 2>&1
cat gpurun_out/c4_sg_debug.log gpurun_out/c4_pytest.log gpurun_out/c4_grade_probe.log; tail -3 gpurun_out/c4_ncu1.log gpurun_out/c4_ncu2.log
