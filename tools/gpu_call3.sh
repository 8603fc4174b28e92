#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for parts in 1 2; do SACB_SG_PARTS=$parts timeout 100 python tools/sg_debug.py 40000 2>&1 | grep "k=4"; done > gpurun_out/c3_sg_debug.log 2>&1
cat gpurun_out/c3_sg_debug.log
