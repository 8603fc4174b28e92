"""Pins oracle/sac_oracle.cpp (reference order, libm) against the golden vectors generated from the reference's own
classes (tests/golden/make_golden.py) and, where oracle/_ref is present, against the reference live. CPU only."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as ol
from helpers import case_planes, case_profile, sha, special_streams
from synth_wav import synth_pcm

REF = (ol.ORDER_REF, ol.MATH_LIBM)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_base_profile(golden):
    vmin, vmax, vdef = ol.base_profile()
    bp = golden["base_profile"]
    assert np.array_equal(vmin, np.array(bp["vmin"], np.float32))
    assert np.array_equal(vmax, np.array(bp["vmax"], np.float32))
    assert np.array_equal(vdef, np.array(bp["vdef"], np.float32))


def test_predict_and_costs_match_reference_goldens(golden):
    vmin, vmax, vdef = ol.base_profile()
    for c in golden["predict"]:
        planes, means, mm = case_planes(c)
        assert [[means[ch], int(mm[2 * ch]), int(mm[2 * ch + 1])] for ch in range(c["nch"])] == c["stats"], c["name"]
        prof = case_profile(c, vmin, vmax, vdef)
        e, rc = ol.oracle_predict(planes, mm, prof, c["k"], c["frm"], c["n"], *REF)
        assert rc == 0
        assert [ol.fnv1a64(x) for x in e] == c["fnv"], c["name"]          # bit-exact residuals
        for kind in (ol.COST_L1, ol.COST_RMS, ol.COST_GOLOMB, ol.COST_BITPLANE):
            got = [ol.oracle_cost(kind, x, math=ol.MATH_LIBM) for x in e]
            assert got == c["costs"][str(kind)], (c["name"], kind)
        got = [ol.oracle_cost(ol.COST_ENTROPY, x, math=ol.MATH_LIBM) for x in e]
        assert np.allclose(got, c["costs"]["2"], rtol=1e-13, atol=0), c["name"]


def test_baseline_md_goldens_mono10():
    """mono10.wav, whole file, default profile: FNV of residuals, entropy and bitplane cost"""
    vmin, vmax, vdef = ol.base_profile()
    pcm = synth_pcm(10, 1, 1).astype(np.int32)
    planes, means, mm = ol.analyse([pcm[:, 0]])
    assert means == [-2]
    e, _ = ol.oracle_predict(planes, mm, vdef, 4, 0, None, *REF)
    assert ol.fnv1a64(e[0]) == "a30578e84f209c5b"
    assert abs(ol.oracle_cost(ol.COST_ENTROPY, e[0], math=ol.MATH_LIBM) - 619005.976981) < 1e-5
    assert abs(ol.oracle_cost(ol.COST_L1, e[0]) - 465.905116) < 1e-6
    assert ol.oracle_cost(ol.COST_BITPLANE, e[0], math=ol.MATH_LIBM) == 619917.0


def test_bitplane_payloads_match_reference(golden):
    vmin, vmax, vdef = ol.base_profile()
    by_name = {b["name"]: b for b in golden["bitplane"]}
    for c in golden["predict"]:
        planes, means, mm = case_planes(c)
        e, _ = ol.oracle_predict(planes, mm, case_profile(c, vmin, vmax, vdef), c["k"], c["frm"], c["n"], *REF)
        payload, maxbpn = ol.oracle_bitplane_encode(ol.s2u(e[0]), math=ol.MATH_LIBM)
        g = by_name[c["name"]]
        assert (maxbpn, len(payload), sha(payload)) == (g["maxbpn"], g["nbytes"], g["sha1"]), c["name"]
        assert np.array_equal(ol.oracle_bitplane_decode(payload, len(e[0]), maxbpn, math=ol.MATH_LIBM), e[0])
    for name, u in special_streams().items():
        g = by_name["special_" + name]
        payload, maxbpn = ol.oracle_bitplane_encode(u, math=ol.MATH_LIBM)
        assert (maxbpn, len(payload), sha(payload)) == (g["maxbpn"], g["nbytes"], g["sha1"]), name


def test_canonical_math_gives_the_reference_bitplane_bytes(golden):
    """the coder is integer except PredictLaplace/table construction: canonical math must not change a byte"""
    by_name = {b["name"]: b for b in golden["bitplane"]}
    for name, u in special_streams().items():
        payload, maxbpn = ol.oracle_bitplane_encode(u, math=ol.MATH_CANON)
        g = by_name["special_" + name]
        assert (len(payload), sha(payload)) == (g["nbytes"], g["sha1"]), name
    lib = ol.oracle()
    for avg, bpn, want in golden["laplace"]:
        for m in (ol.MATH_LIBM, ol.MATH_CANON):
            lib.saco_set_modes(ol.ORDER_REF, m)
            assert lib.saco_predict_laplace(avg, bpn) == want


def test_dds_matches_reference_sequence(golden):
    vmin, vmax, vdef = ol.base_profile()
    idx = [i for i in range(58) if i not in (56, 57)]
    xmin = vmin[idx].astype(np.float64); xmax = vmax[idx].astype(np.float64); xs = vdef[idx].astype(np.float64)
    lib = ol.oracle()
    for g in golden["dds"]:
        trace = []

        def cb(xp, n, _u):
            x = np.ctypeslib.as_array(xp, shape=(n,)).copy()
            trace.append(x)
            z = (x - xmin) / (xmax - xmin)
            return float(np.sum((z - 0.37) ** 2) + 0.05 * np.sum(np.cos(9 * z)))

        fn = ol.COST_CB(cb)
        xb = np.zeros(56)
        fb = lib.saco_dds_run(56, ol._p(xmin, ol._f64p), ol._p(xmax, ol._f64p), ol._p(xs, ol._f64p), g["nfunc"], g["num_threads"],
                              g["sigma"], fn, None, ol._p(xb, ol._f64p))
        assert len(trace) == g["evals"]
        t = np.stack(trace)
        assert sha(t[np.lexsort(t.T[::-1])]) == g["trace_sha1"]
        assert fb == g["best"] and sha(xb) == g["xbest_sha1"]


def test_frame_records_match_reference(golden):
    """FrameCoder::Predict()+Encode() of the reference vs saco_encode_frame: profile, stats and payload bytes"""
    lib = ol.oracle()
    lib.saco_set_modes(*REF)
    _, _, vdef = ol.base_profile()
    for g in golden["frame"]:
        pcm = synth_pcm(g["secs"], g["nch"], g["seed"]).astype(np.int32)
        kw = g["cfg"]
        cfg = (C.c_int * 8)(kw.get("optimize", 0), int(round(kw.get("fraction", 0) * 1e6)), kw.get("maxnfunc", 0),
                            kw.get("num_threads", 0), int(round(kw.get("sigma", 0.2) * 1e6)), 4, kw.get("cost_kind", 2), 20 * 44100)
        prof = vdef.copy()
        s = [np.ascontiguousarray(pcm[:, ch]) for ch in range(g["nch"])]
        out = np.zeros(8 * len(s[0]) + 4096, np.uint8)
        nb = lib.saco_encode_frame(g["nch"], len(s[0]), ol._p(s[0], ol._i32p), ol._p(s[1], ol._i32p) if g["nch"] > 1 else None,
                                   ol._p(prof, ol._f32p), cfg, ol._p(out, ol._u8p), len(out))
        rec = out[:nb]
        assert sha(prof) == g["profile_sha1"], g["name"]
        pos = 4 + 58 * 4
        assert int.from_bytes(rec[:4].tobytes(), "little") == len(s[0])
        assert np.array_equal(np.frombuffer(rec[4:pos].tobytes(), np.float32), prof)
        for ch in range(g["nch"]):
            bs = int.from_bytes(rec[pos:pos + 4].tobytes(), "little")
            mean, mn, mx = np.frombuffer(rec[pos + 4:pos + 16].tobytes(), "<i4")
            assert [int(mean), int(mn), int(mx), int(rec[pos + 16])] == g["stats"][ch][:4], g["name"]
            assert bs == g["payload_len"][ch] and sha(rec[pos + 18:pos + 18 + bs]) == g["payload_sha1"][ch], g["name"]
            pos += 18 + bs
        assert pos == nb
        # and the restated decoder inverts it
        d = [np.zeros(len(s[0]), np.int32) for _ in range(g["nch"])]
        n_out = C.c_int(0)
        used = lib.saco_decode_frame(g["nch"], ol._p(rec, ol._u8p), nb, ol._p(d[0], ol._i32p), ol._p(d[1], ol._i32p) if g["nch"] > 1 else None,
                                     C.byref(n_out))
        assert used == nb and all(np.array_equal(a, b) for a, b in zip(d, s))


def test_live_reference_if_present():
    """when oracle/_ref travels with the repo: random profiles, oracle(ref order) == reference, bit for bit"""
    ref = ol.ref_lib(nc=True)
    if ref is None:
        pytest.skip("oracle/_ref not built")
    vmin, vmax, vdef = ol.base_profile()
    rng = np.random.default_rng(4242)
    pcm = synth_pcm(1, 2, 77).astype(np.int32)
    planes, means, mm = ol.analyse([pcm[:, 0], pcm[:, 1]])
    rf = ol.RefFrame(ref, 2, 20 * 44100)
    rf.set_samples([pcm[:, 0], pcm[:, 1]]); rf.analyse()
    for trial in range(3):
        u = rng.random(58).astype(np.float32)
        prof = (vmin + u * (vmax - vmin)).astype(np.float32)
        prof[28] = min(prof[28], 1500); prof[31] = min(prof[31], 1200)
        er = rf.predict_window(prof, 200, 8000, trial % 2 == 0)
        eo, _ = ol.oracle_predict(planes, mm, prof, 4 if trial % 2 == 0 else 1, 200, 8000, *REF)
        assert all(np.array_equal(a, b) for a, b in zip(er, eo))


def test_b200_order_is_statistically_the_reference(golden):
    """canonical (B200) summation order and elementary functions change rounding only: costs within 1e-3 relative"""
    vmin, vmax, vdef = ol.base_profile()
    for c in golden["predict"][:4]:
        planes, means, mm = case_planes(c)
        prof = case_profile(c, vmin, vmax, vdef)
        e, rc = ol.oracle_predict(planes, mm, prof, c["k"], c["frm"], c["n"], ol.ORDER_B200, ol.MATH_CANON)
        assert rc == 0
        for ch in range(c["nch"]):
            l1 = ol.oracle_cost(ol.COST_L1, e[ch])
            assert abs(l1 - c["costs"]["0"][ch]) <= 1e-3 * c["costs"]["0"][ch], c["name"]
        # round trip through the restated decoder in the same arithmetic
        if c["k"] == 1:
            s = ol.oracle_unpredict(e, mm, prof, ol.ORDER_B200, ol.MATH_CANON)
            assert all(np.array_equal(s[ch][:c["n"]], planes[ch][c["frm"]:c["frm"] + c["n"]]) for ch in range(c["nch"])) or c["frm"] != 0


def _sparse_golden():
    import json
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden_sparse import sparse_pcm
    return json.load(open(os.path.join(ROOT, "tests", "golden", "golden_sparse.json"))), sparse_pcm


def test_sparse_pcm_map_matches_reference():
    """Remap::Analyse/Map/Unmap and MapEncoder (map.cpp) vs the restatement: coded map bytes, rank-mapped residuals, and
    the decoded map equals the analysed one (tests/golden/make_golden_sparse.py)"""
    g, sparse_pcm = _sparse_golden()
    lib = ol.oracle()
    lib.saco_set_modes(*REF)
    for c in g["map"]:
        raw = np.ascontiguousarray(sparse_pcm(c["kind"], c["secs"], c["nch"], c["seed"])[:, 0])
        buf = np.zeros(1 << 16, np.uint8)
        nb = lib.saco_map_encode(ol._p(raw, ol._i32p), len(raw), ol._p(buf, ol._u8p), len(buf))
        assert nb == c["nbytes"] and sha(buf[:nb]) == c["sha1"], c["kind"]
        ul = np.zeros(32769, np.uint8); uh = np.zeros(32769, np.uint8)
        lib.saco_map_decode(ol._p(buf, ol._u8p), nb, ol._p(ul, ol._u8p), ol._p(uh, ol._u8p))
        pos = np.unique(raw[(raw > 0)]); neg = np.unique(-raw[raw < 0])
        wl = np.zeros(32769, np.uint8); wh = np.zeros(32769, np.uint8)
        wh[pos[pos <= 32768]] = 1; wl[neg[neg <= 32768]] = 1
        assert np.array_equal(ul, wl) and np.array_equal(uh, wh), c["kind"]
        rng = np.random.default_rng(c["seed"])
        n = c["n"]
        pred = rng.integers(-33000, 33000, n).astype(np.int32)
        err = (rng.laplace(0, 90, n)).astype(np.int32)
        m = np.zeros(n, np.int32); um = np.zeros(n, np.int32)
        lib.saco_remap(ol._p(raw, ol._i32p), len(raw), ol._p(pred, ol._i32p), ol._p(err, ol._i32p), n, 0, ol._p(m, ol._i32p))
        lib.saco_remap(ol._p(raw, ol._i32p), len(raw), ol._p(pred, ol._i32p), ol._p(m, ol._i32p), n, 1, ol._p(um, ol._i32p))
        assert sha(m) == c["map_sha1"] and sha(um) == c["unmap_sha1"], c["kind"]


def test_sparse_pcm_frame_records_match_reference():
    """FrameCoder::EncodeMonoFrame with sparse_pcm=1 (the CLI default): which channels are coded rank-mapped, the block
    header flag (1<<9 | maxbpn_map) and the payload bytes; the restated decoder (MapEncoder::Decode + Unmap) inverts them"""
    g, sparse_pcm = _sparse_golden()
    lib = ol.oracle()
    lib.saco_set_modes(*REF)
    _, _, vdef = ol.base_profile()
    for c in g["frame"]:
        pcm = sparse_pcm(c["kind"], c["secs"], c["nch"], c["seed"])
        s = [np.ascontiguousarray(pcm[:, ch]) for ch in range(c["nch"])]
        cfg = (C.c_int * 8)(0, 0, 0, 0, 200000, 4, 2, 20 * 44100)
        prof = vdef.copy()
        out = np.zeros(8 * len(s[0]) + 4096, np.uint8)
        mapped = (C.c_int * 2)(0, 0)
        nb = lib.saco_encode_frame2(c["nch"], len(s[0]), ol._p(s[0], ol._i32p), ol._p(s[1], ol._i32p) if c["nch"] > 1 else None,
                                    ol._p(prof, ol._f32p), cfg, 1, ol._p(out, ol._u8p), len(out), mapped)
        rec = out[:nb]
        pos = 4 + 58 * 4
        for ch in range(c["nch"]):
            st = c["stats"][ch]
            bs = int.from_bytes(rec[pos:pos + 4].tobytes(), "little")
            mean, mn, mx = np.frombuffer(rec[pos + 4:pos + 16].tobytes(), "<i4")
            flag = int(rec[pos + 16]) | (int(rec[pos + 17]) << 8)
            assert [int(mean), int(mn), int(mx)] == st[:3] and mapped[ch] == st[5], c["name"]
            assert flag == (((1 << 9) | c["maxbpn_map"][ch]) if st[5] else st[3]), c["name"]
            assert bs == c["payload_len"][ch] and sha(rec[pos + 18:pos + 18 + bs]) == c["payload_sha1"][ch], c["name"]
            pos += 18 + bs
        assert pos == nb
        d = [np.zeros(len(s[0]), np.int32) for _ in range(c["nch"])]
        n_out = C.c_int(0)
        used = lib.saco_decode_frame(c["nch"], ol._p(rec, ol._u8p), nb, ol._p(d[0], ol._i32p), ol._p(d[1], ol._i32p) if c["nch"] > 1 else None,
                                     C.byref(n_out))
        assert used == nb and all(np.array_equal(a, b) for a, b in zip(d, s)), c["name"]
        # with --sparse-pcm=0 nothing is mapped
        nb0 = lib.saco_encode_frame2(c["nch"], len(s[0]), ol._p(s[0], ol._i32p), ol._p(s[1], ol._i32p) if c["nch"] > 1 else None,
                                     ol._p(prof, ol._f32p), cfg, 0, ol._p(out, ol._u8p), len(out), mapped)
        assert list(mapped)[:c["nch"]] == [0] * c["nch"] and (nb0 > nb) == any(x[5] for x in c["stats"])


def test_file_level_frame_plan_and_mapped_blocks_match_the_reference_cli():
    """whole files with sparse stretches: the host frame plan + the restatement's per-block mapped/unmapped choice equal what
    the UNMODIFIED reference CLI wrote (frame lengths and `sparse_pcm:` flags of --listfull; tests/golden/make_golden_split.py)"""
    import io, json, sys, wave
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden_split import case_pcm
    from helpers import oracle_file_image
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_split.json")))
    for name in ("mono_sparse_middle", "stereo_sparse_tail"):
        c = [x for x in g["cases"] if x["name"] == name][0]
        pcm = case_pcm(c["nch"], c["sr"], c["secs"], c["seed"], [tuple(s) for s in c["sparse"]])
        b = io.BytesIO()
        with wave.open(b, "wb") as w:
            w.setnchannels(c["nch"]); w.setsampwidth(2); w.setframerate(c["sr"]); w.writeframes(pcm.astype("<i2").tobytes())
        img, flags = oracle_file_image(b.getvalue(), REF)
        assert flags == c["sparse_pcm_flags"], (name, flags, c["sparse_pcm_flags"])
        assert img[:4] == b"SAC2" and len(img) < len(b.getvalue())


def test_whole_files_are_byte_identical_to_the_reference_cli():
    """Whole .sac files: host container plan (product, sac_container_plan) + the restatement's frame records (reference
    order, libm) == the file the UNMODIFIED reference CLI (built with -ffp-contract=off) wrote, byte for byte: adaptive
    sub-frame split, rank-mapped blocks, --sparse-pcm=0, DDS searches (sequential and population) warm-started from frame to
    frame (tests/golden/make_golden_files.py)."""
    import hashlib, json, sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden_files import FILES, wav_of
    from helpers import oracle_file_image
    g = {f["name"]: f for f in json.load(open(os.path.join(ROOT, "tests", "golden", "golden_files.json")))["files"]}
    assert len(g) == len(FILES) >= 6
    for f in FILES:
        wav = wav_of(f["src"])
        assert hashlib.sha1(wav).hexdigest() == g[f["name"]]["wav_sha1"]
        img, _ = oracle_file_image(wav, REF, sparse=f.get("sparse", 1), optimize=f["optimize"])
        from helpers import reference_view
        assert len(img) == g[f["name"]]["sac_len"] and hashlib.sha1(reference_view(img)).hexdigest() == g[f["name"]]["sac_sha1"], f["name"]
        # The canonical arithmetic the GPU path computes in (its own summation order and exp/log/pow)
        # differs from the reference's in the last bits of a prediction; on files this short no rounded prediction flips, and
        # the file is STILL the reference's, byte for byte (on long full-scale files the bytes drift apart by ~0.004 %).
        img2, _ = oracle_file_image(wav, (ol.ORDER_B200, ol.MATH_CANON), sparse=f.get("sparse", 1), optimize=f["optimize"])
        assert hashlib.sha1(reference_view(img2)).hexdigest() == g[f["name"]]["sac_sha1"], f["name"]


def test_frame_search_that_moves_matches_reference():
    """saco_encode_frame with searches that change tens of dimensions (population and sequential DDS, L1 / entropy / bitplane
    objectives): the profile written into the frame record equals the reference FrameCoder's (tests/golden/make_golden_search.py)"""
    import hashlib, json
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_search.json")))
    lib = ol.oracle()
    lib.saco_set_modes(*REF)
    _, _, vdef = ol.base_profile()
    for c in g["cases"]:
        kw = c["cfg"]
        pcm = synth_pcm(c["secs"], c["nch"], c["seed"]).astype(np.int32)
        s = [np.ascontiguousarray(pcm[:, ch]) for ch in range(c["nch"])]
        cfg = (C.c_int * 8)(1, int(round(kw["fraction"] * 1e6)), kw["maxnfunc"], kw["num_threads"], int(round(kw["sigma"] * 1e6)), 4, kw["cost_kind"],
                            20 * 44100)
        prof = vdef.copy()
        out = np.zeros(8 * len(s[0]) + 4096, np.uint8)
        nb = lib.saco_encode_frame(c["nch"], len(s[0]), ol._p(s[0], ol._i32p), ol._p(s[1], ol._i32p) if c["nch"] > 1 else None,
                                   ol._p(prof, ol._f32p), cfg, ol._p(out, ol._u8p), len(out))
        assert nb > 0 and hashlib.sha1(prof.tobytes()).hexdigest() == c["profile_sha1"], c["name"]
        assert np.array_equal(np.frombuffer(out[4:4 + 58 * 4].tobytes(), np.float32), prof)


def test_reference_best_profile_frame_record():
    """the profile the reference CLI's own --best search ended with on the bench stream's first frame (stages of 6348 / 1873 / 335 taps,
    OLS of 20 and 69 regressors: the parameter ranges a real search reaches, far from the default) on 110 250 sample-frames:
    the restatement's record carries the stats, payload lengths and payload bytes of the reference's FrameCoder
    (tests/golden/make_golden_best_profile.py, generated from libsacref_nc.so)"""
    import json
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_best_profile.json")))
    prof = np.frombuffer(bytes.fromhex(g["profile_hex"]), "<f4").copy()
    assert sha(prof) == g["profile_sha1"]
    pcm = synth_pcm(g["secs"], g["nch"], g["seed"]).astype(np.int32)[:g["n"]]
    s = [np.ascontiguousarray(pcm[:, ch]) for ch in range(g["nch"])]
    lib = ol.oracle()
    lib.saco_set_modes(*REF)
    cfg = (C.c_int * 8)(0, 0, 0, 0, 0, 1, ol.COST_BITPLANE, 20 * 44100)
    out = np.zeros(8 * len(s[0]) + 4096, np.uint8)
    p = prof.copy()
    nb = lib.saco_encode_frame(2, len(s[0]), ol._p(s[0], ol._i32p), ol._p(s[1], ol._i32p), ol._p(p, ol._f32p), cfg, ol._p(out, ol._u8p), len(out))
    rec = out[:nb]
    assert np.array_equal(p, prof)
    pos = 4 + 58 * 4
    for ch in range(2):
        bs = int.from_bytes(rec[pos:pos + 4].tobytes(), "little")
        mean, mn, mx = np.frombuffer(rec[pos + 4:pos + 16].tobytes(), "<i4")
        assert [int(mean), int(mn), int(mx), int(rec[pos + 16])] == g["stats"][ch][:4]
        assert bs == g["payload_len"][ch] and sha(rec[pos + 18:pos + 18 + bs]) == g["payload_sha1"][ch], ch
        pos += 18 + bs
    assert pos == nb
