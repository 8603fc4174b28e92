#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T0=$SECONDS
# launch list of the bench command itself, shortened (2 timed frames, 129 search steps each): every kernel of the path with its
# duration, serialised by ncu -- shares, not a bench value
timeout -s KILL 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/bench_launches_r2.csv \
  python bench.py --steps 2 --warmup 3 --nfunc 129 > gpurun_out/c26_bench_under_ncu.json 2> gpurun_out/c26_bench_under_ncu.err
echo "ncu bench wall $((SECONDS-T0)) s rc=$?"
T0=$SECONDS
timeout -s KILL 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_ref.json 2> gpurun_out/c26_ref.err
echo "reference arm wall $((SECONDS-T0)) s rc=$?"
wc -l gpurun_out/bench_launches_r2.csv gpurun_out/bench_r2_ref.json gpurun_out/c26_bench_under_ncu.json
cut -c1-400 gpurun_out/bench_r2_ref.json
