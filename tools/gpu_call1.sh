#!/bin/bash
# round 2, GPU call 1: micro-benchmarks (DMMA / TMA), sanity of the speculative search on the GPU, bitplane streams per CTA,
# search schedules x frames in flight with the round-1 kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
tools/ubench/dmma_tma > gpurun_out/ubench_dmma_tma.txt 2>&1
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -k "frame_record or eval_population or dedup or multi_frame" 2>&1 | tail -5 > gpurun_out/c1_pytest.log
for v in "" _ps4 _ps6 _ps7; do echo "lib$v"; SAC_B200_LIB=$PWD/sac_b200/libsac_b200$v.so timeout 150 python tools/scale_probe.py 20000 128,768 2>&1 | tail -3; done > gpurun_out/c1_streams.log 2>&1
{
timeout 300 python tools/sched_probe.py gen 128 3 3 257
SACB_LPT=1 timeout 300 python tools/sched_probe.py gen 128 3 3 257
timeout 400 python tools/sched_probe.py spec 16 8 8 80
timeout 500 python tools/sched_probe.py spec 16 16 16 80
SACB_LPT=1 timeout 500 python tools/sched_probe.py spec 16 16 16 80
timeout 400 python tools/sched_probe.py spec 32 8 8 80
} > gpurun_out/c1_sched.jsonl 2> gpurun_out/c1_sched.err
tail -3 gpurun_out/c1_pytest.log; cat gpurun_out/c1_streams.log; cat gpurun_out/c1_sched.jsonl; tail -5 gpurun_out/c1_sched.err; cat gpurun_out/ubench_dmma_tma.txt
