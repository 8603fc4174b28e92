#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c7_launches.csv python tools/grade_probe.py 20000 128 > gpurun_out/c7_probe_under_ncu.log 2>&1
SACB_GRADE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:cascade_sg -c 1 -o gpurun_out/casc_sg_r2b -f python tools/prof_run.py 8 6000 > gpurun_out/c7_ncu1.log 2>&1
SACB_GRADE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:ols_warp -c 1 -o gpurun_out/ols_warp_r2b -f python tools/prof_run.py 8 6000 > gpurun_out/c7_ncu2.log 2>&1
python tools/launch_summary.py gpurun_out/c7_launches.csv 2>&1 | tail -20
