"""CPU-only checks of the product's host side: the C ABI loads and exports every declared symbol, the DDS driver
reproduces the reference's candidate sequence, presets, model tables. No compute entry point is called."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_lib as ol
import sac_b200 as sb
from helpers import sha

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "sac_b200.h")).read()
    names = set(re.findall(r"\b(sac_[a-z0-9_]+)\s*\(", hdr))
    names -= {"sac_eval_fn"}
    assert len(names) >= 20
    L = sb.lib()
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing
    assert b"sac_b200" in L.sac_version()


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(sb.SacError):
        sb.Engine(0)
    assert b"no CPU fallback" in sb.lib().sac_last_error()


def test_base_profile_and_presets(golden):
    vmin, vmax, vdef = sb.base_profile()
    bp = golden["base_profile"]
    assert np.array_equal(vdef, np.array(bp["vdef"], np.float32)) and np.array_equal(vmin, np.array(bp["vmin"], np.float32))
    assert np.array_equal(vmax, np.array(bp["vmax"], np.float32))
    c = sb.make_cfg("best")    # cmdline.cpp:145-150
    assert (c.optimize, c.fraction, c.maxnfunc, c.sigma, c.cost_kind, c.optk) == (1, 0.5, 1000, 0.25, sb.COST_BITPLANE, 4)
    c = sb.make_cfg("high")    # cmdline.cpp:129-134
    assert (c.optimize, c.fraction, c.maxnfunc, c.sigma, c.cost_kind) == (1, 0.1, 100, 0.20, sb.COST_ENTROPY)
    c = sb.make_cfg("normal")
    assert c.optimize == 0 and c.max_framelen == 20 and c.zero_mean == 1


def test_dds_driver_matches_reference_goldens(golden):
    """sac_dds_run (product) against the traces recorded from the reference's OptDDS"""
    vmin, vmax, vdef = sb.base_profile()
    idx = sb.SEARCH_DIMS
    xmin = vmin[idx].astype(np.float64); xmax = vmax[idx].astype(np.float64); xs = vdef[idx].astype(np.float64)
    for g in golden["dds"]:
        trace = []

        def f(X):
            trace.extend(list(X))
            z = (X - xmin) / (xmax - xmin)
            return np.sum((z - 0.37) ** 2, axis=1) + 0.05 * np.sum(np.cos(9 * z), axis=1)

        best, xb = sb.dds_run(f, xmin, xmax, xs, g["nfunc"], g["num_threads"], g["sigma"])
        t = np.stack(trace)
        assert len(t) == g["evals"]
        assert sha(t[np.lexsort(t.T[::-1])]) == g["trace_sha1"]
        assert sha(xb) == g["xbest_sha1"] and abs(best - g["best"]) <= 1e-12 * abs(g["best"])


def test_model_tables_equal_reference_tables():
    st = np.zeros(32768, np.int16); sq = np.zeros(4095, np.int16)
    sb.lib().sac_model_tables(st.ctypes.data_as(C.c_void_p), sq.ctypes.data_as(C.c_void_p))
    fwd = np.zeros(32768, np.int32); inv = np.zeros(4095, np.int32)
    ol.oracle().saco_logdomain_tables(fwd.ctypes.data_as(C.c_void_p), inv.ctypes.data_as(C.c_void_p))
    assert np.array_equal(st.astype(np.int32), fwd) and np.array_equal(sq.astype(np.int32), inv)
    ref = ol.ref_lib()
    if ref is not None:
        ref.ref_logdomain_tables(fwd.ctypes.data_as(C.c_void_p), inv.ctypes.data_as(C.c_void_p))
        assert np.array_equal(st.astype(np.int32), fwd) and np.array_equal(sq.astype(np.int32), inv)


def test_product_does_not_reference_the_oracle():
    """the shipped path must not import, link or execute anything under oracle/"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sac_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "liboracle" not in txt and "sac_oracle" not in txt and "oracle_lib" not in txt, fn
    import subprocess
    out = subprocess.run(["ldd", sb.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_cli_listing_of_a_reference_file_matches_the_reference_cli():
    """container fidelity: `sac --list/--listfull` on a .sac written by the UNMODIFIED reference CLI
    prints exactly what the reference prints for it (header fields, metadata size, MD5 formatting quirk, per-frame block
    headers, coefficient/blocks byte counts). Fixture and expected text: tests/golden/make_cli_fixture.sh. No GPU needed."""
    import subprocess
    gold = os.path.join(ROOT, "tests", "golden")
    cli = os.path.join(ROOT, "sac_b200", "sac")
    assert os.path.exists(cli), "build the CLI first (make -C sac_b200/csrc)"
    # second file: a frame the reference coded rank-mapped (sparse PCM): block flag 1<<9 | maxbpn_map
    for stem, flag, ext in (("ref_stereo_quarter_normal", "--listfull", "listfull"), ("ref_stereo_quarter_normal", "--list", "list"),
                            ("ref_mono_sparse_normal", "--listfull", "listfull")):
        out = subprocess.run([cli, flag, stem + ".sac"], cwd=gold, capture_output=True, text=True, timeout=60).stdout
        got = out[out.index("Open:"):].splitlines()
        want = open(os.path.join(gold, "%s.%s.txt" % (stem, ext))).read().splitlines()
        assert got == want, (flag, got, want)
    assert "sparse_pcm: 1" in open(os.path.join(gold, "ref_mono_sparse_normal.listfull.txt")).read()


def test_de_and_cma_drivers_match_reference_goldens():
    """sac_de_run (product, --opt-cfg=de) against traces recorded from the reference's OptDE (tests/golden/make_golden_de.py):
    every evaluated vector in order, the incumbent and its cost -- including a cost function with ties (std::sort order)
    and nfunc_max below the population size"""
    import hashlib, json
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_de.json")))
    vmin, vmax, vdef = sb.base_profile()
    idx = list(sb.SEARCH_DIMS)
    xmin = vmin[idx].astype(np.float64); xmax = vmax[idx].astype(np.float64); xs = vdef[idx].astype(np.float64)
    sha = lambda a: hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()
    for c in g["cases"]:
        trace = []

        def f(X):
            out = []
            for x in X:
                trace.append(x.copy())
                z = (x - xmin) / (xmax - xmin)
                v = float(np.sum((z - 0.37) ** 2) + 0.05 * np.sum(np.cos(9 * z)))
                out.append(float(np.floor(v * 8)) if c["ties"] else v)
            return out

        best, xb = sb.de_run(f, xmin, xmax, xs, c["nfunc"], c["sigma"])
        assert len(trace) == c["evals"], c
        assert sha(np.stack(trace)) == c["trace_sha1"], c
        assert best == c["best"] and sha(xb) == c["xbest_sha1"], c
    # OptCMA (--opt-cfg=cma): (1+1)-ES, Cholesky of the adapted covariance every step, sigma bootstrapping from 0
    for c in g["cma"]:
        trace = []
        best, xb = sb.cma_run(f, xmin, xmax, xs, c["nfunc"], c["sigma"])
        assert len(trace) == c["evals"], c
        assert sha(np.stack(trace)) == c["trace_sha1"], c
        assert best == c["best"] and sha(xb) == c["xbest_sha1"], c


def test_adaptive_subframe_split_matches_the_reference_cli():
    """Codec::Analyse (adaptive sub-frame split, SURVEY section 8 f-2, host part): for WAVs with sparse stretches the frame
    lengths chosen per 20-s read equal those the UNMODIFIED reference CLI wrote (tests/golden/make_golden_split.py)"""
    import json, sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from make_golden_split import case_pcm
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_split.json")))
    for c in g["cases"]:
        pcm = case_pcm(c["nch"], c["sr"], c["secs"], c["seed"], [tuple(s) for s in c["sparse"]])
        frames, maxf = [], 20 * c["sr"]
        for first in range(0, len(pcm), maxf):
            blk = pcm[first:first + maxf]
            sub = sb.analyse_subframes([blk[:, ch] for ch in range(c["nch"])], c["sr"])
            assert sum(l for _, l, _ in sub) == len(blk) and all(s == sum(x[1] for x in sub[:i]) for i, (s, _, _) in enumerate(sub))
            frames += [l for _, l, _ in sub]
        assert frames == c["frames"], (c["name"], frames, c["frames"])


def test_container_prefix_and_frame_plan_match_the_reference_cli():
    """WAV parsing + .sac container, host only: for WAV images of every shape src/file/wav.cpp handles
    (8/16/24-bit, fmt sizes 16/18/40 incl. WAVE_FORMAT_EXTENSIBLE, chunks before/after 'data', odd chunk sizes, a truncated
    data chunk, several frames) sac_container_plan writes exactly the bytes the UNMODIFIED reference CLI wrote ahead of the
    first frame record (header, metadata, MD5 of the PCM bytes) and plans the same frame records
    (tests/golden/make_golden_container.py)."""
    import hashlib, json, sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from make_golden_container import wav_case
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_container.json")))
    assert len(g["cases"]) >= 11 and any(c.get("rejected") for c in g["cases"])
    for c in g["cases"]:
        wav = wav_case(c["name"])
        assert hashlib.sha1(wav).hexdigest() == c["wav_sha1"], c["name"]
        if c.get("rejected"):                                                    # e.g. stray bytes after the last chunk
            with pytest.raises(sb.SacError):
                sb.container_plan(sb.make_cfg("normal"), wav)
            continue
        prefix, frames, st = sb.container_plan(sb.make_cfg("normal"), wav)
        from helpers import reference_view
        assert reference_view(prefix).hex() == c["prefix_hex"], c["name"]     # the reference's bytes but for the arithmetic-variant byte
        assert frames == c["frames"] and st.nframes == len(frames) and st.numsamples == sum(frames), c["name"]
    # sample containers the codec cannot restore (24 valid bits in 32, a depth of 0 bits): refused, never coded as silence
    assert len(g["refused_here"]) >= 2
    for c in g["refused_here"]:
        wav = wav_case(c["name"])
        assert hashlib.sha1(wav).hexdigest() == c["wav_sha1"], c["name"]
        with pytest.raises(sb.SacError, match="unsupported input format"):
            sb.container_plan(sb.make_cfg("normal"), wav)
    # inputs the reference CLI refuses (cmdline.cpp:253-262, wav.cpp:196-203) are refused with an error, not coded
    bad = bytearray(wav_case("pcm16_stereo")); bad[20] = 3                       # WAVE_FORMAT_IEEE_FLOAT
    with pytest.raises(sb.SacError):
        sb.container_plan(sb.make_cfg("normal"), bytes(bad))
    with pytest.raises(sb.SacError):
        sb.container_plan(sb.make_cfg("normal"), b"RIFF\x04\x00\x00\x00WAVX")


def test_cli_encode_prints_the_reference_header_then_fails_loudly_without_a_gpu(tmp_path):
    """`sac --encode` console layout follows CmdLine::Process (cmdline.cpp:245-262): the Open line and PrintWav come from the
    host container plan; without a CUDA device the tool then stops with an error and exit code 1 (no CPU fallback)."""
    import subprocess, sys
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden_container import wav_case
    wav = wav_case("pcm24_stereo_extensible")
    (tmp_path / "a.wav").write_bytes(wav)
    r = subprocess.run([os.path.join(ROOT, "sac_b200", "sac"), "--encode", "--best", "a.wav", "a.sac"], cwd=tmp_path, capture_output=True, text=True, timeout=60)
    lines = r.stdout.splitlines()
    i = lines.index("Open: 'a.wav': ok (%d Bytes)" % len(wav))
    assert lines[i + 1:i + 4] == ["  WAVE  Codec: PCM (2304 kbps)", "  48000Hz 24 Bit  Stereo", "  9600 Samples [00:00:00.200]"]   # as the reference prints
    assert r.returncode == 1 and "no CPU fallback" in r.stderr and not (tmp_path / "a.sac").exists()


def test_recorded_bench_lines_follow_the_contract():
    """the bench lines kept under profiles/ carry every key of the bench.py contract (metric of BASELINE.json, roofline,
    cpu_baseline, e2e with the copied bytes, clocks, launch count)"""
    import json
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    for fn in ("bench_r1_n1.json", "bench_r1_n2_nfunc129.json", "bench_r2_n1.json", "bench_r2_n1_default.json", "bench_r2_n2_nfunc129.json"):
        d = json.load(open(os.path.join(ROOT, "profiles", fn)))
        assert d["metric"].split(";")[0] in base["metric"] and d["unit"] == "MSamples/s" and d["higher_is_better"] is True
        for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config", "clocks", "gpu_launches"):
            assert k in d, (fn, k)
        assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
        assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["value"] > 0
        assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} and d["e2e"]["h2d_bytes_per_step"] > 0
        assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
        assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if d["n_gpus"] == 1:
            assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] in ("reference", "port")


def test_reference_arm_prints_one_json_line_and_nothing_else():
    """`bench.py --impl reference` (the CPU arm; the one place outside tests/ that may run oracle/): stdout is ONE JSON line with the
    contract's keys although the reference's classes print diagnostics with printf; a non-zero rank under torchrun prints nothing"""
    import json
    import subprocess
    import sys
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.split("\n") if l.strip()]
    assert len(lines) == 1, r.stdout[:2000]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "MSamples/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["extrapolated"] is True and d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "MSamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "recorded steps of the GPU arm" in d["cpu_baseline"]["sample"]          # timed on the candidates the GPU arm evaluated
    assert "reference_concurrency" in d["config"] and d["cpu_baseline"]["single_stream_value"] < d["value"]
    env["RANK"] = "1"; env["WORLD_SIZE"] = "2"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_dds_first_generation_travels_with_the_start_vector_and_nothing_else_changes():
    """population DDS: the product sends the start vector and the first generation as ONE batch (run_mt's first generation
    does not depend on f(xstart), dds.cpp:66-83). Evaluated points IN ORDER, incumbent and cost equal the restatement's
    OptDDS (which evaluates them one after the other as the reference does), also for degenerate budgets."""
    vmin, vmax, vdef = sb.base_profile()
    idx = list(sb.SEARCH_DIMS)
    xmin = vmin[idx].astype(np.float64); xmax = vmax[idx].astype(np.float64); xs = vdef[idx].astype(np.float64)
    D = len(xs)
    lib = ol.oracle()

    def obj(x):
        z = (x - xmin) / (xmax - xmin)
        return float(np.sum((z - 0.41) ** 2) + 0.07 * np.sum(np.cos(11 * z)))

    for nfunc, nt, sigma in ((1, 4, 0.25), (2, 4, 0.25), (5, 8, 0.2), (33, 8, 0.25), (64, 64, 0.25), (130, 128, 0.25), (50, 0, 0.2)):
        otrace = []

        def cb(xp, n, _u):
            x = np.ctypeslib.as_array(xp, shape=(n,)).copy()
            otrace.append(x)
            return obj(x)

        fn = ol.COST_CB(cb)
        oxb = np.zeros(D)
        ofb = lib.saco_dds_run(D, ol._p(xmin, ol._f64p), ol._p(xmax, ol._f64p), ol._p(xs, ol._f64p), nfunc, nt, sigma, fn, None, ol._p(oxb, ol._f64p))
        trace, batches = [], []

        def f(X):
            batches.append(len(X))
            trace.extend(list(X))
            return [obj(x) for x in X]

        best, xb = sb.dds_run(f, xmin, xmax, xs, nfunc, nt, sigma)
        assert len(trace) == len(otrace) == max(nfunc, 1) and all(np.array_equal(a, b) for a, b in zip(trace, otrace)), (nfunc, nt)
        assert best == ofb and np.array_equal(xb, oxb), (nfunc, nt)
        if nt > 0:
            assert batches[0] == 1 + min(nfunc - 1, nt) and sum(batches) == nfunc      # one launch set fewer than generations + 1
        else:
            assert batches == [1] * nfunc


def test_product_search_driver_reaches_the_reference_frame_profile():
    """FrameCoder::Optimize end to end on the CPU: the PRODUCT's DDS driver (population: start vector + first generation as one
    batch; sequential) fed with the restatement's objective (PredictFrame k=4 + cost on the centred window, reference order
    and libm) ends with exactly the 58 floats the reference's FrameCoder::Predict() ended with (tens of dimensions moved;
    tests/golden/make_golden_search.py)."""
    import hashlib, json
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_search.json")))
    vmin, vmax, vdef = ol.base_profile()
    idx = list(sb.SEARCH_DIMS)
    xmin = vmin[idx].astype(np.float64); xmax = vmax[idx].astype(np.float64); xs = vdef[idx].astype(np.float64)
    from synth_wav import synth_pcm
    for c in g["cases"]:
        assert c["changed"] >= 20
        kw = c["cfg"]
        pcm = synth_pcm(c["secs"], c["nch"], c["seed"]).astype(np.int32)
        planes, means, mm = ol.analyse([np.ascontiguousarray(pcm[:, ch]) for ch in range(c["nch"])])
        n = len(pcm); nopt = min(n, int(np.ceil(20 * 44100 * kw["fraction"]))); start = (n - nopt) // 2      # libsac.cpp:367-371

        def f(X):
            out = []
            for x in X:
                prof = vdef.copy(); prof[idx] = x.astype(np.float32)                                      # libsac.cpp:394
                e, _ = ol.oracle_predict(planes, mm, prof, 4, start, nopt, ol.ORDER_REF, ol.MATH_LIBM)
                out.append(sum(ol.oracle_cost(kw["cost_kind"], ee, ol.MATH_LIBM) for ee in e))
            return out

        best, xb = sb.dds_run(f, xmin, xmax, xs, kw["maxnfunc"], kw["num_threads"], kw["sigma"])
        prof = vdef.copy(); prof[idx] = xb.astype(np.float32)                                             # libsac.cpp:418-420
        assert hashlib.sha1(prof.tobytes()).hexdigest() == c["profile_sha1"], c["name"]


def test_profile_mapping_equals_reference_setparam():
    """FrameCoder::SetParam as the product maps a profile for its kernels (sac_profile_params, host only)
    against the reference's own SetParam for the default profile, random profiles in the search box, swapped channel order
    and box corners: all 59 values bit for bit (tests/golden/make_golden_params.py)."""
    import json, sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden_params import profiles
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_params.json")))["cases"]
    vmin, vmax, vdef = sb.base_profile()
    ps = profiles(vmin, vmax, vdef)
    assert len(ps) == len(g) == 16
    for p, want in zip(ps, g):
        got = sb.profile_params(p)
        assert len(got) == len(want) == 59
        assert [float(x).hex() for x in got] == want


def test_frame_prologue_stats_equal_the_reference():
    """mean / min / max of a channel as the block header carries them (FrameCoder::AnalyseMonoChannel + zero-mean,
    SURVEY section 8 a12): product host code against the reference's FrameCoder on awkward inputs -- negative and fractional
    means (floor), constants, one sample, full-scale, with and without --zero-mean. Live against oracle/_ref when it is there,
    and always against the formula the restatement uses (itself pinned through whole frame records)."""
    rng = np.random.default_rng(7)
    cases = [np.array([5], np.int32), np.array([-5], np.int32), np.array([-1, -2], np.int32), np.array([1, 2], np.int32),
             np.full(1000, -32768, np.int32), np.full(999, 32767, np.int32), np.zeros(17, np.int32),
             rng.integers(-32768, 32768, 4001).astype(np.int32), (rng.integers(-300, 100, 777) - 7).astype(np.int32),
             rng.integers(-(1 << 23), 1 << 23, 513).astype(np.int32)]
    ref = ol.ref_lib(nc=True)
    for s in cases:
        for zm in (1, 0):
            got = sb.frame_stats(s, zm)
            mean = int(np.floor(int(s.astype(np.int64).sum()) / float(len(s)))) if zm else 0
            assert got == [mean, int(s.min()) - mean, int(s.max()) - mean]
            if ref is not None:
                rf = ol.RefFrame(ref, 1, 1 << 24, zero_mean=zm)
                rf.set_samples([s]); rf.analyse()
                assert rf.stats(0)[:3] == got, (len(s), zm)


def test_speculative_sequential_search_equals_run_single():
    """sac_dds_run_spec: run_single in speculative batches -- accepted sequence, best cost and best vector equal the one-at-a-time
    search for every batch width (the reference's sequential traces are pinned by the goldens above)"""
    vmin, vmax, vdef = sb.base_profile()
    idx = sb.SEARCH_DIMS
    xmin, xmax, xs = vmin[idx].astype(float), vmax[idx].astype(float), vdef[idx].astype(float)
    rng = np.random.default_rng(7)
    tgt = xmin + (xmax - xmin) * rng.random(56)
    wts = rng.random(56) + 0.1

    def f(X):
        Z = (X - tgt) / (xmax - xmin)
        return (wts * Z * Z).sum(1) + 0.05 * np.abs(np.sin(37 * Z)).sum(1)

    for nf, sigma in ((1, 0.2), (2, 0.2), (60, 0.25), (400, 0.2)):
        seen = []

        def g(X):
            seen.append(X.copy()); return f(X)

        b0, x0 = sb.dds_run(g, xmin, xmax, xs, nf, 0, sigma)
        seq = np.concatenate(seen)
        assert len(seq) == nf
        for spec in (1, 2, 5, 16, 64):
            batches = []

            def h(X):
                batches.append(X.copy()); return f(X)

            b1, x1, ev = sb.dds_run_spec(h, xmin, xmax, xs, nf, sigma, spec)
            assert b1 == b0 and np.array_equal(x1, x0), (nf, spec)
            assert ev >= nf and ev == sum(len(b) for b in batches) and max(len(b) for b in batches) <= max(spec, 1) + 1
            if spec == 1:
                assert ev == nf and np.array_equal(np.concatenate(batches), seq)


def test_speculative_search_with_ties_nan_and_constant_costs():
    """costs that never improve, improve on every step, or are not finite: the walk over a batch stops at the FIRST success only"""
    xmin, xmax, xs = np.zeros(5), np.ones(5), np.full(5, 0.5)
    for fn in (lambda X: np.ones(len(X)), lambda X: X.sum(1), lambda X: np.where(X[:, 0] > 0.6, np.nan, X.sum(1)), lambda X: -np.arange(len(X), dtype=float) - X.sum(1)):
        for nf in (7, 80):
            b0, x0 = sb.dds_run(fn, xmin, xmax, xs, nf, 0, 0.2)
            for spec in (3, 16):
                if fn.__code__.co_consts and "arange" in fn.__code__.co_names:
                    continue                                         # cost depends on the position in the batch: not a function of x
                b1, x1, ev = sb.dds_run_spec(fn, xmin, xmax, xs, nf, 0.2, spec)
                assert (b1 == b0 or (np.isnan(b1) and np.isnan(b0))) and np.array_equal(x1, x0), (nf, spec)


def test_batch_entry_point_validates_its_arguments_without_a_gpu(tmp_path):
    """sac_encode_files / sac_encode_file_multi (several GPUs from the C++ host): argument errors are reported before any device
    work -- null engines, and two inputs that would be written to the same output path"""
    import ctypes as C
    L = sb.lib()
    cfg = sb.make_cfg("normal")
    st = sb.FileStats()
    none = (C.c_void_p * 1)(None)
    assert L.sac_encode_file_multi(none, 1, C.byref(cfg), b"a.wav", b"a.sac", C.byref(st)) == -3      # SAC_E_ARG
    assert L.sac_encode_files(none, 1, C.byref(cfg), 0, None, None, None, None) == -3
    fake = (C.c_void_p * 1)(C.c_void_p(1))                        # never dereferenced: the duplicate check comes first
    wp = (C.c_char_p * 2)(b"x/a.wav", b"y/a.wav"); sp = (C.c_char_p * 2)(b"out/a.sac", b"out/a.sac")
    assert L.sac_encode_files(fake, 1, C.byref(cfg), 2, wp, sp, None, None) == -3
    assert b"same output" in L.sac_last_error()
    sp2 = (C.c_char_p * 2)(b"out/a.sac", None)
    assert L.sac_encode_files(fake, 1, C.byref(cfg), 2, wp, sp2, None, None) == -3 and b"null path" in L.sac_last_error()
