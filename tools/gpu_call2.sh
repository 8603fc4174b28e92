#!/bin/bash
# round 2, GPU call 2: first light of the search-grade kernels (parity vs canonical + timing)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_grade.py -x -q 2>&1 | tail -25 > gpurun_out/c2_pytest.log
timeout 200 python tools/grade_probe.py 60000 128 > gpurun_out/c2_grade_probe.log 2>&1
timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -k "corrupt or roundtrip_with_metadata" 2>&1 | tail -5 > gpurun_out/c2_pytest2.log
cat gpurun_out/c2_pytest.log gpurun_out/c2_grade_probe.log gpurun_out/c2_pytest2.log
