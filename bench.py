#!/usr/bin/env python
"""bench.py -- encode MSamples/s at --best (stereo 16-bit 44.1 kHz), BASELINE.json's metric.

Workload (config.workload): BASELINE.json configs[2] -- the 60-s stereo synthetic WAV (seed 3; every rank its own copy),
`--best --opt-reset --opt-cfg=dds,128`: per 20-s frame a DDS search of 1000 evaluations (the reference's population variant
OptDDS::run_mt in generations of 128) of the OLS+NLMS predictor over a 441 000-sample window scored by the real bitplane
coder, then the final k=1 pass and the payload. The search evaluations run on the search-grade kernels (same formulas,
free summation order and schedule); the final pass and the bitstream are canonical. One STEP = one 20-s frame (882 000
stereo sample-frames; the codec's own unit of work, FrameCoder::Predict+Encode); the K timed steps cycle through the stream's
three frames and are in flight together on the GPU (frame_parallel = 2: a host thread + stream per frame, as a batch of
files or an --opt-reset file is encoded). A sample is one PCM sample-frame (the tool's numsamples).
`--gen 0` runs the reference's DEFAULT sequential search instead (speculative batches, identical trajectory): exact, and a
chain of ~100 dependent batches per frame -- several times slower on this path.

  value = e2e   ONE timed leg through sac_frames_encode with pinned HOST planes: the H2D copy of the planes (7 MB per step,
                0.01 % of a step) and the D2H of the payload are inside. A separate device-resident leg would only repeat the
                number (round 1 measured both: they differed by noise) and double the wall time.
  roofline / fp64   dominant kernel class by CUDA-event time over the timed region; algorithmic bytes and flops of one
                default-profile generation over its device time; traffic from the committed ncu capture
  cpu_baseline      the reference's own classes (oracle/_ref, built from /root/reference) on the host cores, bounded sample

Warm-up steps run the same path (every kernel class, engines, streams, pools) on 2-s excerpts with one short generation, so
that the driver's `--steps 20 --warmup 5` fits its time limit; the timed steps are full frames with the full search.
`--impl reference` times the reference CPU implementation alone (same metric / config), see reference_arm().
Multi-GPU (torchrun, one rank per GPU): weak scaling, every rank encodes its own copy of the stream (same seed: equal work per GPU, nothing shared or cached across ranks); no data-path
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

REAL_STDOUT = 1

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))

SR = 44100
FRAME = 20 * SR
METRIC = "encode MSamples/s at --best (stereo 16-bit 44.1kHz)"
CAND_FIXTURE = os.path.join(ROOT, "tests", "golden", "bench_candidates.json")
TRAFFIC_FIXTURE = os.path.join(ROOT, "profiles", "ncu_traffic.json")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--gen", type=int, default=int(os.environ.get("SAC_BENCH_GEN", "128")),
                    help="N > 0 = --opt-cfg=dds,N (run_mt, the reference's population search); 0 = its sequential search in speculative batches")
    ap.add_argument("--spec", type=int, default=int(os.environ.get("SAC_BENCH_SPEC", "16")), help="candidates per speculative batch")
    ap.add_argument("--grade", type=int, default=int(os.environ.get("SAC_BENCH_GRADE", "1")), help="1 = search-grade kernels for the search")
    ap.add_argument("--nfunc", type=int, default=1000, help="DDS evaluations per frame (--best: 1000)")
    ap.add_argument("--seconds", type=int, default=60)
    ap.add_argument("--inflight", type=int, default=int(os.environ.get("SAC_BENCH_INFLIGHT", "32")), help="frames in flight per GPU (at most the timed steps)")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """nvidia-smi clocks and throttle reasons of one GPU DURING the timed region (rank 0's GPU: the line is rank 0's). One query every
    few seconds: a query takes the driver's lock, and the timed region lasts minutes"""
    def __init__(self, index, period=3.0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.period = index, [], False, period

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
            time.sleep(self.period)

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


def load_candidates():
    """steps of the sequential --best search of the bench stream's first frame as the GPU arm evaluated them (every 40th
    of the 1000, recorded by tools/record_bench_candidates.py on a B200): the CPU is timed on the same work"""
    try:
        with open(CAND_FIXTURE) as f:
            return json.load(f)
    except Exception:
        return None


class StdoutToStderr:
    """the reference's own code prints diagnostics with printf (model/domain.h); this script's stdout carries ONE JSON line, so
    file descriptor 1 points at stderr while the CPU arm runs"""
    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
    def __exit__(self, *a):
        os.dup2(self.saved, 1)
        os.close(self.saved)


def cpu_reference_sample(frames, cores, nevals, nfunc, budget_s=None):
    with StdoutToStderr():
        return cpu_reference_sample_inner(frames, cores, nevals, nfunc)


def cpu_reference_sample_inner(frames, cores, nevals, nfunc):
    """the reference's own objective (PredictFrame k=4 + CostBitplane per channel, libsac.cpp:389-397) on the --best window of
    frame 0 for `nevals` candidates of the real search trajectory, `cores` of them concurrently (one single-threaded
    evaluation per core: the arrangement that wastes nothing, i.e. `cores` frames or files searched side by side), plus one
    final pass + encode on part of a frame; extrapolated linearly in the evaluation count (src/opt/dds.cpp:44)."""
    import oracle_lib as ol
    ref = ol.ref_lib(nc=False)
    kind = "reference" if ref is not None else "port"
    planes, means, mm = ol.analyse(frames[0])
    vmin, vmax, vdef = ol.base_profile()
    n, frm = 441000, 220500
    fixture = load_candidates()
    dims = [i for i in range(58) if i not in (56, 57)]
    profs = []
    if fixture:
        xs = fixture["x"]
        step = max(1, len(xs) // nevals)
        for x in xs[step // 2::step][:nevals]:                 # spread over the whole search
            p = vdef.copy(); p[dims] = np.asarray(x, np.float32); profs.append(p)
        src = "%d of the %d recorded steps of the GPU arm's own search of this frame (tests/golden/bench_candidates.json)" % (len(profs), len(xs))
    else:
        rng = np.random.default_rng(0)
        for i in range(nevals):
            p = vdef.copy()
            if i:
                for j in rng.integers(0, 56, 6):
                    if j not in (9, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 37, 38, 41, 45):
                        p[j] = np.float32(np.clip(p[j] * (1 + 0.05 * rng.standard_normal()), vmin[j], vmax[j]))
            profs.append(p)
        src = "%d default-order candidates (no recorded trajectory available: the CPU is under-costed by about 2.5x)" % len(profs)

    def one_eval(prof):
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
        res = list(ex.map(one_eval, profs))
    wall = time.perf_counter() - t0
    t_single = float(np.mean([r[0] for r in res]))         # seconds of one core per evaluation, `cores` running side by side
    # final pass (k=1 over 882 000) + payload: measured on an eighth of a frame, scaled
    t = time.perf_counter()
    q = FRAME // 8
    if ref is not None:
        rf = ol.RefFrame(ref, 2, FRAME)
        rf.set_samples([p[:q] for p in frames[0]])
        rf.predict(); rf.encode()
    else:
        pl, _, m2 = ol.analyse([p[:q] for p in frames[0]])
        e, _ = ol.oracle_predict(pl, m2, vdef, 1, 0, q, ol.ORDER_REF, ol.MATH_LIBM)
        for x in e:
            ol.oracle_bitplane_encode(ol.s2u(x), math=ol.MATH_LIBM)
    t_final = (time.perf_counter() - t) * 8
    # `cores` searches side by side, each one core: a frame takes nfunc * t_single + t_final of one core
    t_frame_core = nfunc * t_single + t_final
    value = cores * FRAME / t_frame_core / 1e6
    single_stream = FRAME / (nfunc * t_single / 2.0 + t_final / 2.0) / 1e6   # stock CLI: one frame at a time, the two channels on two threads (mt_mode = 2)
    return {"value": value, "unit": "MSamples/s", "cores": cores, "kind": kind, "extrapolated": True,
            "sample": "%s, %d at a time, one core each (%.2f s per evaluation of the 441000-sample stereo window: PredictFrame k=4 + "
                      "CostBitplane) + final pass on an eighth of a frame (%.1f s per frame); extrapolated to %d evaluations + final pass "
                      "per 882000-sample frame and to %d searches side by side (the arrangement without wasted evaluations)"
                      % (src, cores, t_single, t_final, nfunc, cores),
            "seconds_per_eval_core": t_single, "seconds_final": t_final, "wall_s": wall,
            "single_stream_value": single_stream,
            "single_stream_note": "the stock CLI (num_threads = 0, mt_mode = 2): one frame at a time, two threads"}


def workload_config(args, nfr=3):
    """the `config` object both arms print (the reference arm times the same workload on the host cores)"""
    sched = ("sequential DDS (run_single/SSC0, the reference's default) in speculative batches of <= %d" % args.spec) if args.gen <= 0 else \
            ("--opt-cfg=dds,%d (run_mt/SSC1), generations of %d" % (args.gen, args.gen))
    return {"workload": "configs[2]: stereo 16-bit 44.1kHz 60s synthetic WAV (seed 3, one copy per rank), --best --opt-reset; step = one 20-s frame "
                        "(882000 sample-frames): %d DDS steps, %s, window 441000, CostBitplane, k=4, search-grade kernels = %d; final pass k=1 + "
                        "bitplane payload (canonical); %d frames in flight per GPU (one host thread + stream set each)" % (args.nfunc, sched, args.grade, max(1, min(args.inflight, max(args.steps, 1)))),
            "search": "dds_sequential_speculative" if args.gen <= 0 else "dds_population", "spec": args.spec, "generation": args.gen, "grade": args.grade,
            "nfunc": args.nfunc, "frames": nfr, "frames_in_flight": max(1, min(args.inflight, max(args.steps, 1))),
            "l2": "inputs per step (7 MB planes + 3.5 MB of p_lpc and 1.7 MB of residuals per chain, hundreds of chains in flight) "
                  "exceed the 126 MB L2; no flush"}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    frames = stream_frames(20, 3)
    vals = []
    cb = None
    t_all = time.perf_counter()
    for i in range(args.warmup + args.steps):
        cb = cpu_reference_sample(frames, cores, nevals=cores, nfunc=args.nfunc)
        if i >= args.warmup:
            vals.append(cb["value"])
        # a step of this arm is a bounded sample; keep the whole run within a few minutes
        if time.perf_counter() - t_all + cb["wall_s"] * 1.5 > 300:
            if not vals:
                vals = [cb["value"]]
            break
    v = float(np.mean(vals))
    cb["value"] = v
    cfg = workload_config(args)
    cfg["reference_concurrency"] = "%d single-threaded evaluations side by side on %d host cores; value extrapolated from %d evaluations per sample" % (cores, cores, cores)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "MSamples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": FRAME / v / 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "extrapolated": True, "samples_taken": len(vals),
            "config": cfg,
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "extrapolated", "single_stream_value", "single_stream_note")},
            "e2e": {"value": v, "unit": "MSamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


def split_records(rec, count):
    """concatenated frame records (u32 numsamples, 58 f32, per channel 18 B header + payload) -> list of bytes"""
    out, pos = [], 0
    for _ in range(count):
        p0 = pos; pos += 4 + 58 * 4
        for ch in range(2):
            nb = int(np.frombuffer(rec[pos:pos + 4].tobytes(), "<u4")[0]); pos += 18 + nb
        out.append(rec[p0:pos].tobytes())
    return out


def emit(line):
    """the ONE line of this script's stdout"""
    os.write(REAL_STDOUT, (json.dumps(line) + "\n").encode())


def main():
    global REAL_STDOUT
    args = parse()
    # libraries under this script print to stdout (NCCL's version banner, the reference's printf diagnostics): descriptor 1 points at
    # stderr for the whole run and the result line goes to the saved descriptor
    sys.stdout.flush()
    REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
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
    frames = stream_frames(args.seconds, 3)          # the same stream on every rank: equal work per GPU (weak scaling), encoded independently
    nfr = len(frames)
    inflight = max(1, min(args.inflight, max(args.steps, 1)))
    cfg = sb.make_cfg("best", num_threads=max(args.gen, 0), spec=args.spec, maxnfunc=args.nfunc, frame_parallel=2, inflight=inflight, reset=1,
                      grade=args.grade)
    # warm-up steps: the same frames, kernels, pools and helper engines with a short search (every kernel class runs)
    warm = sb.make_cfg("best", num_threads=max(args.gen, 0), spec=args.spec, maxnfunc=min(args.nfunc, 24), frame_parallel=2, inflight=inflight, reset=1,
                       grade=args.grade)
    pinned = [[torch.from_numpy(p.copy()).pin_memory() for p in fr] for fr in frames]
    pinned_np = [[t.numpy() for t in fr] for fr in pinned]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(c, count, first):
        fs = [(first + i) % nfr for i in range(count)]
        rec, _ = eng.frames_encode(c, [pinned_np[f] for f in fs], FRAME, None)
        return fs, split_records(rec, count)

    if args.warmup > 0:                                                   # 2-s excerpts: kernels, engines, streams and pools, not the work
        short = [[p[:2 * SR].copy() for p in fr] for fr in frames]
        eng.frames_encode(warm, [short[i % nfr] for i in range(args.warmup)], FRAME, None)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dd_before = eng.dedup_totals(); l_before = eng.launches; tm_before, calls_before = eng.total_timing(); gs_before = eng.grade_stats()
    t0 = time.perf_counter()
    fs, recs = run_steps(cfg, args.steps, args.warmup)
    barrier()
    t_val = time.perf_counter() - t0
    sampler.stop_flag = True
    l_timed = eng.launches - l_before
    dd_timed = [a - b for a, b in zip(eng.dedup_totals(), dd_before)]
    tm, calls = eng.total_timing(); tm_timed = [a - b for a, b in zip(tm, tm_before)]; calls -= calls_before
    gs_timed = [a - b for a, b in zip(eng.grade_stats(), gs_before)]
    out_bytes = {}
    for f, r in zip(fs, recs):
        out_bytes[f] = r
    # kernel-class timing of one representative generation for the roofline (CUDA events on the engine's stream, device alone)
    import oracle_lib as ol   # analyse() only (mean / min / max on the host, numpy) -- outside the timed region
    _, _, vdef = sb.base_profile()
    P = 128
    x0 = np.tile(vdef[sb.SEARCH_DIMS].astype(np.float64), (P, 1))
    pl, mn, mm = ol.analyse(frames[0])
    win = eng.window(pl, mm)
    prev = eng.set_dedup(0)       # the probe wants P identical default chains actually evaluated
    prevg = eng.set_grade(args.grade)
    eng.eval_population(win, 220500, 20000, vdef, x0, sb.COST_BITPLANE, 4)        # pools of the main engine
    barrier()
    eng.eval_population(win, 220500, 441000, vdef, x0, sb.COST_BITPLANE, 4)
    ms, ln = eng.last_timing()
    eng.set_dedup(prev); eng.set_grade(prevg)
    win.close()
    fp64_peak = eng.fp64_peak_gflops()
    t_val = shard.max_over_ranks(t_val, dev)
    value = FRAME * args.steps * world / t_val / 1e6
    # bitstream gather (outside the timed region): the only collective of the path. Unit f*world+rank = frame f of rank's stream
    gathered = None
    if world > 1:
        n_units = world * nfr
        local_units = {f * world + rank: out_bytes.get(f, b"") for f in range(nfr)}
        gathered = shard.gather_bitstreams(local_units, n_units, rank, world, dev)
    if rank == 0:
        chains = 2 * P
        W = 441000
        f0, f1 = algorithmic_flops_per_sample(vdef, 4)
        flops = (f0 + f1) * W * P
        ols_s, casc_s, bp_s = ms[3] * 1e-3, (ms[0] - ms[3]) * 1e-3, ms[1] * 1e-3
        planes_coded = 15                                         # maxbpn + 1 of the synthetic streams' residuals
        kern = {
            # algorithmic HBM bytes per chain-sample: own + other plane in (2 x 4 B), p_lpc out (8 B)
            "ols": (ols_s, chains * W * 16),
            # sample in (4 B), p_lpc in (8 B), residual out (4 B)
            "cascade": (casc_s, chains * W * 16),
            # residual in + S2U-mapped back in place (8 B), then one 4-B read per coded plane, cost out
            "bitplane": (bp_s, chains * W * (8 + 4 * planes_coded) + chains * 8),
        }
        names = {"ols": "ols_warp_kernel" if args.grade else "ols_kernel", "cascade": "cascade_sg_kernel" if args.grade else "cascade_kernel",
                 "bitplane": "bitplane_pipe_kernel"}
        share = {"ols": tm_timed[0], "cascade": tm_timed[1], "bitplane": tm_timed[2]}
        tot = sum(share.values()) or 1.0
        dominant = max(share, key=lambda k: share[k]) if tot > 1.0 else max(kern, key=lambda k: kern[k][0])
        dur, alg_bytes = kern[dominant]
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        traffic = None
        try:
            tj = json.load(open(TRAFFIC_FIXTURE))
            per_chain_sample = tj.get(names[dominant], {}).get("dram_bytes_per_chain_sample")
            if per_chain_sample is not None:
                traffic = per_chain_sample * chains * W
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = alg_bytes / dur / 1e9 if dur > 0 else 0.0
        cores = os.cpu_count() or 1
        if world > 1:             # the contract: on rank 0 at N=1 only (the other arm, --impl reference, is timed at every N)
            cbo = {"value": None, "unit": "MSamples/s", "cores": 0, "kind": "not measured", "sample": "measured at N=1 only"}
        else:
            try:
                cb = cpu_reference_sample(frames, cores, nevals=cores, nfunc=args.nfunc)
                cbo = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "extrapolated", "single_stream_value", "single_stream_note")}
            except Exception as ex:   # the baseline is a reported number, never a dependency of the product path
                cbo = {"value": None, "unit": "MSamples/s", "cores": 0, "kind": "unavailable", "sample": repr(ex)}
        h2d = FRAME * 2 * 4
        d2h = int(np.mean([len(b) for b in recs]))
        pred_s = ols_s + casc_s
        bps = 8.0 * sum(len(b) for b in out_bytes.values()) / (len(out_bytes) * FRAME * 2)
        ref_bytes = {}
        try:
            ref_bytes = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_best_sizes.json")))
        except Exception:
            pass
        f0_ref = ref_bytes.get("bench_frame0", {}).get("frame_record_bytes")
        line = {
            "metric": METRIC, "value": value, "unit": "MSamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_val / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": workload_config(args, nfr),
            "clocks": sampler.summary(),
            "e2e": {"value": value, "unit": "MSamples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "note": "the timed leg IS the end-to-end leg: pinned host planes in, frame records out (H2D + D2H inside)"},
            "gpu_launches": int(l_timed),
            "chains": {"requested": dd_timed[0], "evaluated": dd_timed[1], "ols_stages_evaluated": dd_timed[2],
                       "candidates_evaluated": dd_timed[0] // 2, "search_steps": args.nfunc * args.steps,
                       "search_grade": {"cascade_small": gs_timed[0], "cascade_large": gs_timed[1], "canonical_fallback": gs_timed[2], "jobs_re_evaluated_after_clamp": gs_timed[3]},
                       "note": "speculative batches spend more candidates than search steps; exact de-duplication: chains / OLS stages of a batch with identical inputs run once"},
            "device_ms_by_kernel_class": {"ols": tm_timed[0], "cascade": tm_timed[1], "bitplane": tm_timed[2], "evaluations": calls,
                                          "note": "CUDA events per stream over the timed region; streams overlap, so shares not sums"},
            "roofline": {"bound": "hbm", "kernel": names[dominant], "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic,
                         "note": "dominant kernel class by device time in the timed region; achieved = algorithmic bytes of a generation of %d default "
                                 "chains / its CUDA-event duration (device alone). The path is a set of serial fp64/integer recurrences bound by "
                                 "instruction latency, not by HBM: see `fp64` and DESIGN.md section 5; peak = MEASURED_PEAKS.json hbm_gbs" % chains + ("" if peaks else " (fallback 6650)")},
            "fp64": {"kernel": names["ols"] + " + " + names["cascade"], "achieved_gflops": flops / pred_s / 1e9 if pred_s > 0 else None,
                     "peak_gflops": fp64_peak, "frac": (flops / pred_s / 1e9) / fp64_peak if pred_s > 0 and fp64_peak > 0 else None,
                     "peak_how": "measured in this run: 8 independent DFMA chains/thread, 1184x256 threads, CUDA events",
                     "flops_per_stereo_sample": f0 + f1},
            "kernel_ms_per_generation": {"ols": ms[3], "cascade": ms[0] - ms[3], "bitplane": ms[1], "chains": chains, "window": W,
                                         "profile": "default (one generation of identical default-profile candidates, device alone)"},
            "cpu_baseline": cbo,
            "gathered_bytes": (sum(len(b) for b in gathered) if gathered is not None else None),
            "bytes_per_frame": d2h, "bps": bps,
            "bps_reference": ({"frame0_bytes_ours": len(out_bytes[0]) if 0 in out_bytes else None, "frame0_bytes_reference": f0_ref,
                               "frame0_bps_ours": (8.0 * len(out_bytes[0]) / (FRAME * 2)) if 0 in out_bytes else None,
                               "frame0_bps_reference": 8.0 * f0_ref / (FRAME * 2),
                               "how": ref_bytes.get("bench_frame0", {}).get("how")} if f0_ref else ref_bytes.get("stereo10")),
        }
        emit(line)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
