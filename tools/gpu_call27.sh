#!/bin/bash
# 2 GPUs: the bench as the driver launches it for N=2 (shortened search), then the multi-GPU tests of the C++ host
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T0=$SECONDS
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --steps 2 --warmup 3 --nfunc 129 > gpurun_out/bench_r2_n2_nfunc129.json 2> gpurun_out/c27_bench.err
echo "bench N=2 wall $((SECONDS-T0)) s rc=$?"
T0=$SECONDS
timeout -s KILL 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/c27_ref_n2.json 2> gpurun_out/c27_ref.err
echo "reference arm N=2 wall $((SECONDS-T0)) s rc=$?"
T0=$SECONDS
timeout -s KILL 300 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/c27_multi.log 2>&1
echo "multi tests wall $((SECONDS-T0)) s rc=$?"
tail -3 gpurun_out/c27_multi.log
wc -l gpurun_out/bench_r2_n2_nfunc129.json gpurun_out/c27_ref_n2.json
cut -c1-300 gpurun_out/bench_r2_n2_nfunc129.json
tail -3 gpurun_out/c27_bench.err | cut -c1-300
