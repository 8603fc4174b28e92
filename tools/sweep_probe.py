#!/usr/bin/env python
"""C5 of BASELINE.json: DDS population sweep 64 -> 4096 candidates on a 60-s stereo frame, 1 GPU or N GPUs
(candidates of a generation dealt across the ranks, one all_gather of costs per generation; sac_b200/shard.py).

    python tools/sweep_probe.py --pops 64,256,1024 --window 441000
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/sweep_probe.py

Per population size P: ONE generation of P heterogeneous DDS candidates (first generation around the default profile,
sigma 0.25 -- the expensive kind) over the --best window, timed with CUDA events, max over ranks. Prints one JSON line
per P: candidate evaluations/s and evaluated window samples/s (P x window / seconds). A probe, not the bench."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pops", default="64,128,256,512,1024,2048,4096")
    ap.add_argument("--window", type=int, default=441000)
    ap.add_argument("--sigma", type=float, default=0.25)
    ap.add_argument("--grade", type=int, default=1, help="1 = search-grade kernels (what the bench's search runs on)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import sac_b200 as sb
    from sac_b200 import shard
    from synth_wav import synth_pcm
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = sb.Engine(local)
    eng.set_grade(args.grade)
    pcm = synth_pcm(20, 2, 3).astype(np.int32)
    planes = [np.ascontiguousarray(pcm[:, 0]), np.ascontiguousarray(pcm[:, 1])]
    means = [int(np.floor(p.sum() / len(p))) for p in planes]
    planes = [p - m for p, m in zip(planes, means)]
    mm = [int(planes[0].min()), int(planes[0].max()), int(planes[1].min()), int(planes[1].max())]
    win = eng.window(planes, mm)
    n = min(args.window, len(planes[0])); frm = (len(planes[0]) - n) // 2
    vmin, vmax, vdef = sb.base_profile()
    idx = list(sb.SEARCH_DIMS)
    xmin = vmin[idx].astype(np.float64); xmax = vmax[idx].astype(np.float64); xs = vdef[idx].astype(np.float64)

    def eval_rows(Xs):
        return eng.eval_population(win, frm, n, vdef, Xs, sb.COST_BITPLANE, 4)

    for P in [int(x) for x in args.pops.split(",")]:
        gens = []

        def f(X):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            c = shard.sharded_population_costs(eval_rows, X, rank, world, dev)
            torch.cuda.synchronize()
            gens.append((len(X), shard.max_over_ranks(time.perf_counter() - t0, dev)))
            return c

        # nfunc = 1 + P: the start point, then exactly one generation of P candidates (dds.cpp:63-106)
        sb.dds_run(f, xmin, xmax, xs, 1 + P, P, args.sigma)
        full = [g for g in gens if g[0] >= P]            # the start vector travels with the first generation: P + 1 rows
        if rank == 0 and full:
            sec = full[0][1]
            print(json.dumps({"probe": "population_sweep", "n_gpus": world, "population": P, "window": n, "grade": args.grade, "seconds": round(sec, 3),
                              "evals_per_s": round(full[0][0] / sec, 2), "window_msamples_per_s": round(full[0][0] * n / sec / 1e6, 3)}), flush=True)
    win.close(); eng.close()
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
