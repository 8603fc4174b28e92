#!/bin/bash
# The GPU calls that were left unmeasured when round 1's GPU budget ran out, in the order they pay off. Each line is one
# `gpurun` call (run them one at a time from /root/repo; outputs land in gpurun_out/, copy summaries to profiles/).
# Nothing here is executed by tests or the driver.
set -e
G=/usr/local/graft/bin/gpurun
case "$1" in
  suite)   $G --timeout 700 -- 'timeout 680 python -m pytest tests -m gpu -x -q --durations=10 2>&1 | tail -25' ;;
  bench)   $G --timeout 900 -- 'python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.log; tail -c 3000 gpurun_out/bench_n1.json' ;;
  scale)   for n in 2 4 8; do $G --gpus $n --timeout 400 -- "python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --steps 3 --warmup 3 --nfunc 129 --e2e-steps 1 > gpurun_out/bench_n${n}_nfunc129.json 2> gpurun_out/bench_n${n}.log; tail -c 1500 gpurun_out/bench_n${n}_nfunc129.json"; done ;;
  sweep1)  $G --timeout 600 -- 'timeout 580 python tools/sweep_probe.py --pops 64,128,256,512,1024 > gpurun_out/sweep_n1.jsonl 2>&1; cat gpurun_out/sweep_n1.jsonl' ;;
  sweep8)  $G --gpus 8 --timeout 600 -- 'timeout 580 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 tools/sweep_probe.py --pops 512,1024,2048,4096 > gpurun_out/sweep_n8.jsonl 2>&1; cat gpurun_out/sweep_n8.jsonl' ;;
  batch8)  $G --gpus 8 --timeout 900 -- 'python tools/make_corpus.py /tmp/c4 0.25 && timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29545 -m sac_b200.batch --best --opt-cfg=dds,128 --out /tmp/c4out /tmp/c4/*.wav 2>&1 | tail -5' ;;
  sparse)  $G --timeout 300 -- 'timeout 280 python -m pytest tests/test_gpu_sparse.py -q --durations=5 2>&1 | tail -12' ;;
  streams) # bitplane streams per CTA: rebuild here, measure one default-profile generation + the coder's parity tests on the box
           for n in 4 6 7 2; do touch sac_b200/csrc/bitplane.cu; make -s -C sac_b200/csrc EXTRA=-DSACB_PIPE_STREAMS=$n;
             $G --timeout 300 -- "echo streams=$n; timeout 120 python tools/first_light.py 2>&1 | tail -4; timeout 150 python -m pytest tests/test_gpu_parity.py -q -k 'bitplane or frame_record or eval_population' 2>&1 | tail -3"; done ;;
  lpt)     $G --timeout 400 -- 'for v in 0 1; do echo SACB_LPT=$v; SACB_LPT=$v timeout 170 python tools/gen_probe.py 20000 128 0.67 2>&1 | tail -4; done; SACB_LPT=1 timeout 120 python -m pytest tests/test_gpu_parity.py -q -k "eval_population or frame_record" 2>&1 | tail -3' ;;
  *) echo "usage: $0 suite|bench|scale|sweep1|sweep8|batch8|sparse|streams|lpt" ;;
esac
