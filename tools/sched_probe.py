#!/usr/bin/env python
"""Throughput of the frame encoder under different ways of feeding the GPU (probe, not product): search schedule
(population generations / speculative sequential batches), batch width, frames in flight. Prints one JSON line per run.
usage: sched_probe.py MODE WIDTH INFLIGHT FRAMES NFUNC [GRADE [FRACTION]]   (MODE = gen | spec; FRACTION of the 882000-sample frame
searched, --best = 0.5: the kernels are per-sample recurrences, so time scales linearly with it)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import sac_b200 as sb
from synth_wav import synth_pcm

mode, width, inflight, nframes, nfunc = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
grade = int(sys.argv[6]) if len(sys.argv) > 6 else 0
fraction = float(sys.argv[7]) if len(sys.argv) > 7 else 0.5
FRAME = 20 * 44100
pcm = synth_pcm(60, 2, 3).astype(np.int32)
base = [[np.ascontiguousarray(pcm[f:f + FRAME, 0]), np.ascontiguousarray(pcm[f:f + FRAME, 1])] for f in range(0, len(pcm), FRAME)]
frames = [base[i % len(base)] for i in range(nframes)]
eng = sb.Engine(0)
cfg = sb.make_cfg("best", num_threads=width if mode == "gen" else 0, spec=width, maxnfunc=nfunc, frame_parallel=2 if inflight > 1 else 0,
                  inflight=inflight, reset=1, grade=grade, fraction=fraction, verbose=int(os.environ.get('SACB_VERBOSE', '0')))
warm = sb.make_cfg("best", num_threads=0, spec=4, maxnfunc=3, frame_parallel=2, inflight=inflight, reset=1, grade=grade)
eng.frames_encode(warm, frames[:min(nframes, inflight)], FRAME)          # pools, helper engines
d0 = eng.dedup_totals(); l0 = eng.launches; tm0, c0 = eng.total_timing(); g0 = eng.grade_stats()
t0 = time.perf_counter()
rec, _ = eng.frames_encode(cfg, frames, FRAME)
dt = time.perf_counter() - t0
d1 = eng.dedup_totals(); tm1, c1 = eng.total_timing(); g1 = eng.grade_stats()
print(json.dumps({"fraction": fraction, "mode": mode, "width": width, "inflight": inflight, "frames": nframes, "nfunc": nfunc, "grade": grade, "seconds": round(dt, 2),
                  "candidates": (d1[0] - d0[0]) // 2, "chains_evaluated": d1[1] - d0[1], "ols_evaluated": d1[2] - d0[2],
                  "candidates_per_s": round((d1[0] - d0[0]) / 2 / dt, 2), "s_per_frame": round(dt / nframes, 2), "launches": eng.launches - l0,
                  "bytes": int(len(rec)), "device_ms": [round(a - b, 1) for a, b in zip(tm1, tm0)], "evaluations": c1 - c0, "grade_stats": [a - b for a, b in zip(g1, g0)], "lpt": os.environ.get("SACB_LPT", "0"), "lib": os.path.basename(sb.LIB_PATH)}), flush=True)
eng.close()
