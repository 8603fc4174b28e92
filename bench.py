#!/usr/bin/env python
"""bench.py -- encode MSamples/s at --best (stereo 16-bit 44.1 kHz), BASELINE.json's metric.

Workload (config.workload): BASELINE.json configs[2] -- the 60-s stereo synthetic WAV (seed 3),
`--best`: per 20-s frame a DDS search of 1000 evaluations of the OLS+NLMS predictor over a 441 000-sample window scored
by the real bitplane coder, then the final k=1 pass and the payload. One STEP = one 20-s frame of that stream
(882 000 stereo sample-frames; the codec's own unit of work, FrameCoder::Predict+Encode), steps cycle through the
file's three frames, each frame warm-started from the previous one's optimum as the reference does. A sample is one
PCM sample-frame (the tool's numsamples).

  value  device-resident: the frame's planes are already in HBM when the timed region starts
  e2e    the same step through sac_frames_encode with HOST buffers (H2D of the planes, D2H of the payload inside)
  roofline / fp64   dominant kernel's algorithmic bytes and flops over its CUDA-event duration
  cpu_baseline      the reference's own classes (oracle/_ref, built from /root/reference) on the host cores, bounded sample

`--impl reference` times the reference CPU implementation alone (same metric / config), see reference_arm().
Multi-GPU (torchrun, one rank per GPU): weak scaling, every rank encodes its own stream (seed 3 + rank); no data-path
collective; rank 0 gathers the bitstreams at the end (outside the timed region).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

SR = 44100
FRAME = 20 * SR
METRIC = "encode MSamples/s at --best (stereo 16-bit 44.1kHz)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gen", type=int, default=int(os.environ.get("SAC_BENCH_GEN", "128")), help="DDS generation size (GPU batch per frame)")
    ap.add_argument("--nfunc", type=int, default=1000, help="DDS evaluations per frame (--best: 1000)")
    ap.add_argument("--seconds", type=int, default=60)
    ap.add_argument("--inflight", type=int, default=3, help="frames encoded concurrently per GPU (one stream each); 1 = sequential with warm start")
    ap.add_argument("--e2e-steps", dest="e2e_steps", type=int, default=3, help="steps of the end-to-end leg (host buffers)")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.5)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def stream_frames(seconds, seed):
    from synth_wav import synth_pcm
    pcm = synth_pcm(seconds, 2, seed).astype(np.int32)
    return [[np.ascontiguousarray(pcm[f:f + FRAME, 0]), np.ascontiguousarray(pcm[f:f + FRAME, 1])] for f in range(0, len(pcm), FRAME)]


def algorithmic_flops_per_sample(profile, k):
    """F = F_ols(n,k) + sum_j 10*N_j + F_misc per sample per channel (fma = 2)"""
    def f_ols(n):
        return 2 * n + 2 * n * (n + 1) + 4 * n + (n ** 3 / 2 + 2 * n * n + n) / k
    r = lambda i: int(round(float(profile[i])))
    n0 = r(24) + r(9); n1 = r(25) + r(26) + abs(r(27))
    taps0 = r(28) + r(29) + r(30) + r(37); taps1 = r(31) + r(32) + r(33) + r(38)
    return f_ols(n0) + 10 * taps0 + 300, f_ols(n1) + 10 * taps1 + 300


NFUNC_REF = 1000


def cpu_reference_sample(frames, cores, nevals=None):
    """the reference's own objective (PredictFrame k=4 + CostBitplane per channel, libsac.cpp:389-397) on the --best
    window of frame 0, `cores` candidates concurrently (what --opt-cfg=dds,N does), plus one final pass + encode,
    extrapolated linearly in the evaluation count (src/opt/dds.cpp:44)."""
    import oracle_lib as ol
    ref = ol.ref_lib(nc=False)
    kind = "reference"
    if ref is None:
        kind = "port"
    planes, means, mm = ol.analyse(frames[0])
    vmin, vmax, vdef = ol.base_profile()
    n, frm = 441000, 220500
    nevals = nevals or cores
    rng = np.random.default_rng(0)

    def one_eval(i):
        prof = vdef.copy()
        if i:  # perturbed like a DDS candidate (keeps orders near the default so that the sample is representative)
            idx = rng.integers(0, 56, 6)
            for j in idx:
                if j not in (9, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 37, 38, 41, 45):
                    prof[j] = np.float32(np.clip(prof[j] * (1 + 0.05 * rng.standard_normal()), vmin[j], vmax[j]))
        t = time.perf_counter()
        if ref is not None:
            rf = ol.RefFrame(ref, 2, FRAME)
            rf.set_samples(frames[0]); rf.analyse()
            e = rf.predict_window(prof, frm, n, True)
            c = sum(ref.ref_cost(4, ol._p(x, ol._i32p), n) for x in e)
        else:
            e, _ = ol.oracle_predict(planes, mm, prof, 4, frm, n, ol.ORDER_REF, ol.MATH_LIBM)
            c = sum(ol.oracle_cost(ol.COST_BITPLANE, x, math=ol.MATH_LIBM) for x in e)
        return time.perf_counter() - t, c

    from concurrent.futures import ThreadPoolExecutor
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=cores) as ex:     # ctypes releases the GIL during the call
        res = list(ex.map(one_eval, range(nevals)))
    wall = time.perf_counter() - t0
    per_eval_wall = wall / nevals                          # with `cores` in flight
    t_single = float(np.mean([r[0] for r in res]))
    # final pass (k=1 over 882 000) + payload: measured on a quarter frame, scaled
    t = time.perf_counter()
    q = FRAME // 4
    if ref is not None:
        rf = ol.RefFrame(ref, 2, FRAME)
        rf.set_samples([p[:q] for p in frames[0]])
        rf.predict(); rf.encode()
    else:
        pl, _, m2 = ol.analyse([p[:q] for p in frames[0]])
        e, _ = ol.oracle_predict(pl, m2, vdef, 1, 0, q, ol.ORDER_REF, ol.MATH_LIBM)
        for x in e:
            ol.oracle_bitplane_encode(ol.s2u(x), math=ol.MATH_LIBM)
    t_final = (time.perf_counter() - t) * 4
    t_frame = NFUNC_REF * per_eval_wall + t_final
    return {"value": FRAME / t_frame / 1e6, "unit": "MSamples/s", "cores": cores, "kind": kind,
            "sample": "%d --best objective evaluations (441000-sample stereo window, PredictFrame k=4 + CostBitplane) run %d at a time "
                      "(%.2f s each single-threaded, %.2f s wall per evaluation) + final pass on a quarter frame; extrapolated to %d "
                      "evaluations + final pass per 882000-sample frame" % (nevals, cores, t_single, per_eval_wall, NFUNC_REF),
            "seconds_per_eval": per_eval_wall, "seconds_final": t_final}


def workload_config(args, nfr=3):
    """the `config` object both arms print (the reference arm times the same workload on the host cores)"""
    return {"workload": "configs[2]: stereo 16-bit 44.1kHz 60s synthetic WAV (seed 3+rank), --best --opt-reset; step = one 20-s frame "
                        "(882000 sample-frames): DDS %d evaluations in generations of %d (--opt-cfg=dds,%d: run_mt/SSC1), window 441000, "
                        "CostBitplane, k=4; final pass k=1 + bitplane payload; %d frames in flight per GPU (one stream each)"
                        % (args.nfunc, args.gen, args.gen, args.inflight),
            "generation": args.gen, "nfunc": args.nfunc, "frames": nfr, "frames_in_flight": args.inflight, "e2e_steps": args.e2e_steps,
            "l2": "inputs per step (7 MB planes + 3.5 MB of p_lpc and 1.7 MB of residuals per chain, 256 chains per generation) "
                  "exceed the 126 MB L2; no flush"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    global NFUNC_REF
    NFUNC_REF = args.nfunc
    cores = os.cpu_count() or 1
    frames = stream_frames(20, 3)
    vals = []
    cb = None
    for i in range(args.warmup + args.steps):
        cb = cpu_reference_sample(frames, cores)
        if i >= args.warmup:
            vals.append(cb["value"])
        if i == 0 and args.warmup + args.steps > 1 and cb["seconds_per_eval"] * cores * (args.warmup + args.steps) > 600:
            vals = [cb["value"]]
            break
    v = float(np.mean(vals))
    cb["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "MSamples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": FRAME / v / 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": v, "unit": "MSamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def main():
    global NFUNC_REF
    args = parse()
    NFUNC_REF = args.nfunc
    if args.impl == "reference":
        reference_arm(args)
        return
    import torch
    import torch.distributed as dist
    import sac_b200 as sb
    from sac_b200 import shard
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the sac_b200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = sb.Engine(local)
    import oracle_lib as ol   # analyse() only (mean / min / max on the host, numpy)
    frames = stream_frames(args.seconds, 3 + rank)
    nfr = len(frames)
    # frames of the stream are independent searches (--opt-reset semantics) and run concurrently, one stream each
    cfg = sb.make_cfg("best", num_threads=args.gen, maxnfunc=args.nfunc, frame_parallel=2 if args.inflight > 1 else 0, reset=1)
    # device-resident copies for the `value` leg
    wins, means = [], []
    for fr in frames:
        pl, mn, mm = ol.analyse(fr)
        wins.append(eng.window(pl, mm)); means.append(mn)
    # pinned host planes for the e2e leg
    pinned = [[torch.from_numpy(p.copy()).pin_memory() for p in fr] for fr in frames]
    pinned_np = [[t.numpy() for t in fr] for fr in pinned]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    out_bytes = {}

    def run_steps(count, resident, first):
        """`count` steps = `count` consecutive frames of the stream (cyclic), `inflight` of them per call"""
        s = 0
        while s < count:
            g = min(args.inflight, count - s)
            fs = [(first + s + i) % nfr for i in range(g)]
            if resident:
                rec, _ = eng.frames_encode_resident(cfg, [wins[f] for f in fs], [means[f] for f in fs], FRAME, None)
            else:
                rec, _ = eng.frames_encode(cfg, [pinned_np[f] for f in fs], FRAME, None)
            pos = 0
            for f in fs:            # split the concatenated frame records (u32 numsamples, 58 f32, per channel 18 B header + payload)
                p0 = pos; pos += 4 + 58 * 4
                for ch in range(2):
                    nb = int(np.frombuffer(rec[pos:pos + 4].tobytes(), "<u4")[0]); pos += 18 + nb
                out_bytes[f] = rec[p0:pos].tobytes()
            s += g

    run_steps(args.warmup, True, 0)
    barrier()
    sampler.start()
    dd_before = eng.dedup_totals()
    l_before = eng.launches
    t0 = time.perf_counter()
    run_steps(args.steps, True, args.warmup)
    barrier()
    t_val = time.perf_counter() - t0
    l_timed = eng.launches - l_before
    dd_timed = [a - b for a, b in zip(eng.dedup_totals(), dd_before)]
    # e2e leg: the same steps from pinned HOST planes through sac_frames_encode (H2D of the planes, D2H of the payload inside)
    barrier()
    t0 = time.perf_counter()
    run_steps(args.e2e_steps, False, args.warmup)
    barrier()
    t_e2e = max(time.perf_counter() - t0, 1e-9)
    sampler.stop_flag = True
    # kernel-class timing of one representative generation for the roofline (CUDA events on the engine's stream)
    _, _, vdef = sb.base_profile()
    x0 = np.tile(vdef[sb.SEARCH_DIMS].astype(np.float64), (args.gen, 1))
    prev = eng.set_dedup(0)       # the probe wants `gen` identical default chains actually evaluated
    eng.eval_population(wins[0], 220500, 441000, vdef, x0, sb.COST_BITPLANE, 4)
    barrier()
    eng.eval_population(wins[0], 220500, 441000, vdef, x0, sb.COST_BITPLANE, 4)
    ms, ln = eng.last_timing()
    eng.set_dedup(prev)
    fp64_peak = eng.fp64_peak_gflops()
    t_val = shard.max_over_ranks(t_val, dev); t_e2e = shard.max_over_ranks(t_e2e, dev)
    value = FRAME * args.steps * world / t_val / 1e6
    e2e = FRAME * args.e2e_steps * world / t_e2e / 1e6 if args.e2e_steps > 0 else None   # 0 only for profiling runs
    # bitstream gather (outside the timed region): the only collective of the path. Unit f*world+rank = frame f of rank's stream
    gathered = None
    if world > 1:
        n_units = world * nfr
        local_units = {f * world + rank: out_bytes.get(f, b"") for f in range(nfr)}
        gathered = shard.gather_bitstreams(local_units, n_units, rank, world, dev)
    if rank == 0:
        chains = 2 * args.gen
        W = 441000
        f0, f1 = algorithmic_flops_per_sample(vdef, 4)
        flops = (f0 + f1) * W * args.gen
        ols_s, casc_s, bp_s = ms[3] * 1e-3, (ms[0] - ms[3]) * 1e-3, ms[1] * 1e-3
        planes_coded = 15                                         # maxbpn + 1 of the synthetic streams' residuals
        kern = {
            # algorithmic HBM bytes per chain-sample: own + other plane in (2 x 4 B), p_lpc out (8 B)
            "ols_kernel": (ols_s, chains * W * 16),
            # sample in (4 B), p_lpc in (8 B), residual out (4 B)
            "cascade_kernel": (casc_s, chains * W * 16),
            # residual in + S2U-mapped back in place (8 B), then one 4-B read per coded plane, cost out
            "bitplane_pipe_kernel": (bp_s, chains * W * (8 + 4 * planes_coded) + chains * 8),
        }
        dominant = max(kern, key=lambda k: kern[k][0])
        dur, alg_bytes = kern[dominant]
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = alg_bytes / dur / 1e9 if dur > 0 else 0.0
        cores = os.cpu_count() or 1
        try:
            cb = cpu_reference_sample(frames, min(cores, 8), nevals=min(cores, 8))
            cbo = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as ex:   # the baseline is a reported number, never a dependency of the product path
            cbo = {"value": None, "unit": "MSamples/s", "cores": 0, "kind": "unavailable", "sample": repr(ex)}
        h2d = FRAME * 2 * 4
        d2h = int(np.mean([len(b) for b in out_bytes.values()]))
        pred_s = ols_s + casc_s
        line = {
            "metric": METRIC, "value": value, "unit": "MSamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_val / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": workload_config(args, nfr),
            "clocks": sampler.summary(),
            "e2e": {"value": e2e, "unit": "MSamples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(l_timed),
            "chains": {"requested": dd_timed[0], "evaluated": dd_timed[1], "ols_stages_evaluated": dd_timed[2],
                       "note": "exact de-duplication: chains / OLS stages of a generation with identical inputs run once (DESIGN.md section 4.6)"},
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": None,
                         "note": "the path is a set of serial fp64/integer recurrences bound by instruction latency, not by HBM: "
                                 "see `fp64` and DESIGN.md section 5; peak = MEASURED_PEAKS.json hbm_gbs" + ("" if peaks else " (fallback 6650)")},
            "fp64": {"kernel": "ols_kernel + cascade_kernel", "achieved_gflops": flops / pred_s / 1e9 if pred_s > 0 else None,
                     "peak_gflops": fp64_peak, "frac": (flops / pred_s / 1e9) / fp64_peak if pred_s > 0 and fp64_peak > 0 else None,
                     "peak_how": "measured in this run: 8 independent DFMA chains/thread, 1184x256 threads, CUDA events",
                     "flops_per_stereo_sample": f0 + f1},
            "kernel_ms_per_generation": {"ols": ms[3], "cascade": ms[0] - ms[3], "bitplane": ms[1], "chains": chains, "window": W,
                                         "profile": "default (one generation of identical default-profile candidates, device alone)"},
            "cpu_baseline": cbo,
            "gathered_bytes": (sum(len(b) for b in gathered) if gathered is not None else None),
            "bytes_per_frame": d2h, "bps": 8.0 * sum(len(b) for b in out_bytes.values()) / (len(out_bytes) * FRAME * 2),
        }
        print(json.dumps(line))
    for w in wins:
        w.close()
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
