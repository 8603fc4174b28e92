#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for c in n1 n3 silence dc fullscale ragged; do for k in 1 3 4; do timeout -s KILL 20 python tools/sg_edge.py $c $k 2>&1 | tail -2 | cut -c1-300; [ ${PIPESTATUS[0]} -ne 0 ] && echo "$c $k: rc ${PIPESTATUS[0]}"; done; done > gpurun_out/c9_edge.log 2>&1
cat gpurun_out/c9_edge.log
