#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{ SACB_SG_PARTS=2 timeout 100 python tools/sg_debug.py 40000 2>&1 | grep "k=4"
echo nofma; SAC_B200_LIB=$PWD/sac_b200/libsac_b200_nofma.so SACB_SG_PARTS=2 timeout 100 python tools/sg_debug.py 40000 2>&1 | grep "k=4"; } > gpurun_out/c5_sg_debug.log
cat gpurun_out/c5_sg_debug.log
