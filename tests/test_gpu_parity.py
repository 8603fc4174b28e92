"""-m gpu: the CUDA path through the C ABI against the oracle (canonical arithmetic: bit-exact) and against the golden
vectors recorded from the reference itself (integer paths: bit-exact; fp paths: stated tolerance)."""
import ctypes as C
import io
import struct
import wave

import numpy as np
import pytest

import oracle_lib as ol
import sac_b200 as sb
from helpers import case_planes, case_profile, random_profile, sha, special_streams
from synth_wav import synth_pcm

pytestmark = pytest.mark.gpu
FS = 20 * 44100


def _oracle_frame(nch, planes_raw, cfg_kw, profile):
    lib = ol.oracle()
    lib.saco_set_modes(ol.ORDER_B200, ol.MATH_CANON)
    cfg = (C.c_int * 8)(cfg_kw.get("optimize", 0), int(round(cfg_kw.get("fraction", 0) * 1e6)), cfg_kw.get("maxnfunc", 0),
                        cfg_kw.get("num_threads", 0), int(round(cfg_kw.get("sigma", 0.2) * 1e6)), cfg_kw.get("optk", 4),
                        cfg_kw.get("cost_kind", 2), FS)
    prof = np.ascontiguousarray(profile, np.float32).copy()
    s = [np.ascontiguousarray(p, np.int32) for p in planes_raw]
    out = np.zeros(8 * len(s[0]) + 4096, np.uint8)
    nb = lib.saco_encode_frame(nch, len(s[0]), ol._p(s[0], ol._i32p), ol._p(s[1], ol._i32p) if nch > 1 else None, ol._p(prof, ol._f32p),
                               cfg, ol._p(out, ol._u8p), len(out))
    return out[:nb].copy(), prof


def test_predictor_bit_exact_vs_oracle_on_golden_cases(engine, golden):
    vmin, vmax, vdef = sb.base_profile()
    for c in golden["predict"]:
        planes, means, mm = case_planes(c)
        prof = case_profile(c, vmin, vmax, vdef)
        win = engine.window(planes, mm)
        res, flags = engine.predict(win, [prof], c["frm"], c["n"], c["k"])
        e, rc = ol.oracle_predict(planes, mm, prof, c["k"], c["frm"], c["n"])
        assert flags[0] == rc == 0
        for ch in range(c["nch"]):
            assert np.array_equal(res[0, ch], e[ch]), (c["name"], ch)
            # against the reference's own arithmetic: same statistics (rounding flips only), tolerance 1e-3 on L1
            l1 = float(np.abs(res[0, ch]).mean())
            assert abs(l1 - c["costs"]["0"][ch]) <= 1e-3 * c["costs"]["0"][ch] + 1e-9
        win.close()


def test_predictor_batch_random_profiles_and_hbm_spill(engine):
    """a population with heterogeneous orders, including ones whose tap state exceeds the CTA's shared memory"""
    vmin, vmax, vdef = sb.base_profile()
    rng = np.random.default_rng(31)
    pcm = synth_pcm(1, 2, 8).astype(np.int32)
    planes, means, mm = ol.analyse([pcm[:, 0], pcm[:, 1]])
    win = engine.window(planes, mm)
    profs = [vdef, random_profile(rng, vmin, vmax), random_profile(rng, vmin, vmax, cap0=None, cap1=None), vmax.copy(), vmin.copy()]
    profs[1][27] = -3.0
    n = 3000
    res, flags = engine.predict(win, profs, 100, n, 4)
    for p, prof in enumerate(profs):
        e, rc = ol.oracle_predict(planes, mm, prof, 4, 100, n)
        assert rc == flags[p]
        assert np.array_equal(res[p, 0], e[0]) and np.array_equal(res[p, 1], e[1]), p
    win.close()


def test_predictor_edge_inputs(engine):
    _, _, vdef = sb.base_profile()
    sq = (np.where((np.arange(4000) // 50) % 2 == 0, 32767, -32768)).astype(np.int32)
    for name, planes_raw in (("n1", [np.array([123], np.int32)]), ("silence", [np.zeros(3000, np.int32)] * 2),
                             ("dc", [np.full(2000, -7, np.int32)]), ("fullscale", [sq, -sq]),
                             ("ragged", [synth_pcm(1, 1, 3).astype(np.int32)[:2999, 0]])):
        planes, means, mm = ol.analyse(planes_raw)
        win = engine.window(planes, mm)
        n = len(planes[0])
        for k in (1, 4):
            res, flags = engine.predict(win, [vdef], 0, n, k)
            e, rc = ol.oracle_predict(planes, mm, vdef, k, 0, n)
            assert all(np.array_equal(res[0, ch], e[ch]) for ch in range(len(planes))), (name, k)
        win.close()


def test_costs_vs_oracle_and_reference_goldens(engine, golden):
    vmin, vmax, vdef = sb.base_profile()
    c = golden["predict"][2]
    planes, means, mm = case_planes(c)
    e, _ = ol.oracle_predict(planes, mm, vdef, c["k"], c["frm"], c["n"], ol.ORDER_REF, ol.MATH_LIBM)   # the reference's residuals
    bufs = np.stack(e)
    for kind in (sb.COST_L1, sb.COST_RMS, sb.COST_GOLOMB, sb.COST_BITPLANE):
        assert engine.cost(kind, bufs).tolist() == c["costs"][str(kind)], kind        # exact vs the reference
    got = engine.cost(sb.COST_ENTROPY, bufs)
    assert np.allclose(got, c["costs"]["2"], rtol=1e-12, atol=0)                        # fp: summation order only
    # empty-ish and degenerate buffers
    assert engine.cost(sb.COST_BITPLANE, np.zeros((1, 64), np.int32))[0] == ol.oracle_cost(ol.COST_BITPLANE, np.zeros(64, np.int32))
    assert engine.cost(sb.COST_ENTROPY, np.zeros((1, 64), np.int32))[0] == 0.0


def test_bitplane_payload_bytes_equal_the_reference(engine, golden):
    by_name = {b["name"]: b for b in golden["bitplane"]}
    for name, u in special_streams().items():
        # signed residuals whose S2U image is u
        e = np.where(u % 2 == 1, (u + 1) // 2, -(u // 2)).astype(np.int32)
        assert np.array_equal(ol.s2u(e), u)
        payload, maxbpn = engine.bitplane_encode(e)
        g = by_name["special_" + name]
        assert (maxbpn, len(payload), sha(payload)) == (g["maxbpn"], g["nbytes"], g["sha1"]), name
        assert np.array_equal(engine.bitplane_decode(payload, len(e), maxbpn), e), name
    vmin, vmax, vdef = sb.base_profile()
    for c in golden["predict"][:3]:
        planes, means, mm = case_planes(c)
        e, _ = ol.oracle_predict(planes, mm, case_profile(c, vmin, vmax, vdef), c["k"], c["frm"], c["n"], ol.ORDER_REF, ol.MATH_LIBM)
        payload, maxbpn = engine.bitplane_encode(e[0])
        g = by_name[c["name"]]
        assert (maxbpn, len(payload), sha(payload)) == (g["maxbpn"], g["nbytes"], g["sha1"]), c["name"]
        assert np.array_equal(engine.bitplane_decode(payload, len(e[0]), maxbpn), e[0])


def test_eval_population_equals_oracle_costs(engine):
    vmin, vmax, vdef = sb.base_profile()
    rng = np.random.default_rng(5)
    pcm = synth_pcm(1, 2, 12).astype(np.int32)
    planes, means, mm = ol.analyse([pcm[:, 0], pcm[:, 1]])
    win = engine.window(planes, mm)
    idx = sb.SEARCH_DIMS
    X = np.stack([vdef[idx].astype(np.float64)] + [(vmin[idx] + rng.random(56) * (np.minimum(vmax[idx], 1500) - vmin[idx])).astype(np.float64)
                                                     for _ in range(5)])
    frm, n = 2000, 6000
    for kind in (sb.COST_L1, sb.COST_ENTROPY, sb.COST_BITPLANE):
        got = engine.eval_population(win, frm, n, vdef, X, kind, 4)
        for p in range(len(X)):
            prof = vdef.copy(); prof[idx] = X[p].astype(np.float32)
            e, rc = ol.oracle_predict(planes, mm, prof, 4, frm, n)
            want = sum(ol.oracle_cost(kind, e[ch]) for ch in range(2))
            if kind == sb.COST_ENTROPY:
                assert abs(got[p] - want) <= 1e-10 * want
            else:
                assert got[p] == want, (kind, p)
    win.close()


@pytest.mark.parametrize("nch,kw", [(1, dict(optimize=0)), (2, dict(optimize=0)),
                                    (2, dict(optimize=1, fraction=0.004, maxnfunc=9, sigma=0.2, cost_kind=4)),
                                    (2, dict(optimize=1, fraction=0.004, maxnfunc=13, num_threads=4, sigma=0.25, cost_kind=0))])
def test_frame_record_identical_to_oracle_and_roundtrip(engine, nch, kw):
    """whole frame: analysis, DDS search (sequential and population), final pass, payload, serialisation"""
    pcm = synth_pcm(0.5, nch, 40 + nch).astype(np.int32)
    raw = [pcm[:, ch] for ch in range(nch)]
    _, _, vdef = sb.base_profile()
    cfg = sb.make_cfg(None, **kw)
    rec, prof = engine.frames_encode(cfg, [raw], FS)
    orec, oprof = _oracle_frame(nch, raw, kw, vdef)
    assert np.array_equal(prof, oprof)
    assert len(rec) == len(orec) and np.array_equal(rec, orec)
    dec, used = engine.frame_decode(nch, rec, FS)
    assert used == len(rec) and all(np.array_equal(dec[ch], raw[ch]) for ch in range(nch))
    # cross-decoding with the oracle's decoder (same canonical arithmetic)
    lib = ol.oracle(); lib.saco_set_modes(ol.ORDER_B200, ol.MATH_CANON)
    d = [np.zeros(len(raw[0]), np.int32) for _ in range(nch)]
    n_out = C.c_int(0)
    lib.saco_decode_frame(nch, ol._p(rec, ol._u8p), len(rec), ol._p(d[0], ol._i32p), ol._p(d[1], ol._i32p) if nch > 1 else None, C.byref(n_out))
    assert all(np.array_equal(d[ch], raw[ch]) for ch in range(nch))


def test_frame_size_within_tolerance_of_the_reference(engine, golden):
    """compressed size vs the reference's own fp64 build: stated tolerance 0.1 % per channel payload (rounding flips)"""
    for g in golden["frame"][:2]:
        pcm = synth_pcm(g["secs"], g["nch"], g["seed"]).astype(np.int32)
        rec, _ = engine.frames_encode(sb.make_cfg("normal"), [[pcm[:, ch] for ch in range(g["nch"])]], FS)
        pos = 4 + 58 * 4
        for ch in range(g["nch"]):
            bs = int.from_bytes(rec[pos:pos + 4].tobytes(), "little")
            assert abs(bs - g["payload_len"][ch]) <= 1e-3 * g["payload_len"][ch] + 2, (g["name"], ch, bs, g["payload_len"][ch])
            pos += 18 + bs


def _wav_bytes(pcm, sr=44100, extra_chunks=(), sampwidth=2):
    nch = pcm.shape[1]
    if sampwidth == 2:
        data = pcm.astype("<i2").tobytes()
    elif sampwidth == 1:
        data = (pcm + 128).astype(np.uint8).tobytes()
    else:
        b = pcm.astype("<i4").tobytes()
        data = b"".join(b[i:i + 3] for i in range(0, len(b), 4))
    fmt = struct.pack("<HHIIHH", 1, nch, sr, sr * nch * sampwidth, nch * sampwidth, 8 * sampwidth)
    body = b"WAVE" + b"fmt " + struct.pack("<I", 16) + fmt
    for cid, payload in extra_chunks[:1]:
        body += cid + struct.pack("<I", len(payload)) + payload + (b"\0" if len(payload) & 1 else b"")
    body += b"data" + struct.pack("<I", len(data)) + data + (b"\0" if len(data) & 1 else b"")
    for cid, payload in extra_chunks[1:]:
        body += cid + struct.pack("<I", len(payload)) + payload + (b"\0" if len(payload) & 1 else b"")
    return b"RIFF" + struct.pack("<I", len(body)) + body


@pytest.mark.parametrize("nch,sampwidth,n,sr", [(2, 2, 30011, 8000), (1, 2, 4000, 44100), (1, 1, 5001, 44100), (2, 3, 3000, 48000)])
def test_file_roundtrip_with_metadata(engine, nch, sampwidth, n, sr):
    pcm = synth_pcm(1, nch, 60 + nch).astype(np.int32)[:n]
    if sampwidth == 1:
        pcm = pcm >> 8
    elif sampwidth == 3:
        pcm = pcm * 200 + 17
    wav = _wav_bytes(pcm, sr=sr, extra_chunks=[(b"LIST", b"INFOabc"), (b"id3 ", b"x" * 11)], sampwidth=sampwidth)
    sac, st = engine.encode_memory(sb.make_cfg("normal", max_framelen=1), wav)
    assert sac[:4] == b"SAC2" and st.numsamples == n and st.nch == nch
    assert st.nframes == -(-n // sr)                       # ragged last frame
    back, st2 = engine.decode_memory(sac, len(wav) + 1024)
    assert st2.md5_ok == 1
    assert back == wav


def test_multi_frame_profile_chain_and_frame_parallel(engine):
    """frames of one stream: sequential warm start (reference semantics, final pass overlapped with the next search) and
    the two frame-parallel extensions (1: shared launches, 2: one stream + host thread per frame) all decode; the
    frame-parallel modes search every frame from the base profile, so they agree with each other byte for byte"""
    pcm = synth_pcm(1.5, 2, 77).astype(np.int32)
    frames = [[pcm[:22050, 0], pcm[:22050, 1]], [pcm[22050:44100, 0], pcm[22050:44100, 1]], [pcm[44100:, 0], pcm[44100:, 1]]]
    recs = {}
    for fp in (0, 1, 2):
        cfg = sb.make_cfg(None, optimize=1, fraction=0.005, maxnfunc=6, num_threads=5, sigma=0.2, cost_kind=2, frame_parallel=fp, max_framelen=1)
        rec, prof = engine.frames_encode(cfg, frames, 44100)
        recs[fp] = rec.tobytes()
        pos = 0
        for fr in frames:
            dec, used = engine.frame_decode(2, rec[pos:], 44100)
            assert all(np.array_equal(dec[ch], fr[ch]) for ch in range(2))
            pos += used
        assert pos == len(rec)
    assert recs[1] == recs[2]


def test_dedup_of_identical_chains_is_exact(engine):
    """chains with identical inputs (a candidate that leaves one channel, or only its OLS parameters, untouched) are
    evaluated once: costs and residuals equal those of the run without de-duplication, and fewer chains are launched"""
    vmin, vmax, vdef = sb.base_profile()
    pcm = synth_pcm(0.5, 2, 21).astype(np.int32)
    planes, means, mm = ol.analyse([pcm[:, 0], pcm[:, 1]])
    win = engine.window(planes, mm)
    idx = list(sb.SEARCH_DIMS)
    base = vdef[idx].astype(np.float64)
    X = np.tile(base, (6, 1))
    X[1, idx.index(2)] *= 1.07           # ch0 NLMS mu only: ch1 chain identical to candidate 0, OLS stage of ch0 shared
    X[2, idx.index(14)] *= 0.93          # ch1 NLMS mu only
    X[3, idx.index(0)] = 0.9985          # ch0 OLS lambda: new OLS stage for ch0
    X[4] = X[1]                          # exact duplicate candidate
    X[5, idx.index(41)] = 3.0            # RLS order: both channels change, OLS stages shared
    frm, n = 300, 6000
    prev = engine.set_dedup(1)
    t0 = engine.dedup_totals()
    c_on = engine.eval_population(win, frm, n, vdef, X, sb.COST_BITPLANE, 4)
    t1 = engine.dedup_totals()
    engine.set_dedup(0)
    c_off = engine.eval_population(win, frm, n, vdef, X, sb.COST_BITPLANE, 4)
    t2 = engine.dedup_totals()
    assert np.array_equal(c_on, c_off)
    assert c_on[1] == c_on[4]
    req, ev, ols = [a - b for a, b in zip(t1, t0)]
    assert req == 12 and ev == 7 and ols == 3, (req, ev, ols)   # ch0 chains {0,2},{1,4},{3},{5}; ch1 chains {0,1,3,4},{2},{5}; OLS: ch0 base, ch0 lambda, ch1
    assert [a - b for a, b in zip(t2, t1)] == [12, 12, 12]
    # residual scatter with duplicates
    profs = []
    for x in X:
        p = vdef.copy(); p[idx] = x.astype(np.float32); profs.append(p)
    engine.set_dedup(1)
    r_on, _ = engine.predict(win, profs, frm, n, 4)
    engine.set_dedup(0)
    r_off, _ = engine.predict(win, profs, frm, n, 4)
    engine.set_dedup(prev)
    assert np.array_equal(r_on, r_off)
    win.close()


def test_frame_encode_with_de_search_roundtrips(engine):
    """--opt-cfg=de / cma through the frame seam: DE generations (start vector, 29 initial samples, trial vectors) evaluated
    as populations, CMA one candidate per step; the records decode bit-exactly and are not larger than the unoptimised one"""
    pcm = synth_pcm(0.5, 2, 31).astype(np.int32)
    raw = [pcm[:, 0], pcm[:, 1]]
    base, _ = engine.frames_encode(sb.make_cfg("normal", max_framelen=1), [raw], 44100)
    cfg = sb.make_cfg(None, optimize=1, fraction=0.25, maxnfunc=36, sigma=0.2, cost_kind=sb.COST_BITPLANE, max_framelen=1, search=sb.SEARCH_DE)
    l0 = engine.launches
    rec, prof = engine.frames_encode(cfg, [raw], 44100)
    # 3 population evaluations + the final pass, 3 kernels each; the final pass also runs the 4 sparse-PCM kernels (sparse.cu)
    assert engine.launches - l0 == 3 * 3 + 3 + 4
    dec, used = engine.frame_decode(2, rec, 44100)
    assert used == len(rec) and np.array_equal(dec[0], raw[0]) and np.array_equal(dec[1], raw[1])
    assert len(rec) <= len(base)
    rec2, _ = engine.frames_encode(sb.make_cfg(None, optimize=1, fraction=0.25, maxnfunc=6, cost_kind=sb.COST_ENTROPY, max_framelen=1,
                                               search=sb.SEARCH_CMA), [raw], 44100)
    dec, used = engine.frame_decode(2, rec2, 44100)
    assert used == len(rec2) and np.array_equal(dec[0], raw[0]) and np.array_equal(dec[1], raw[1])
    with pytest.raises(sb.SacError):
        engine.frames_encode(sb.make_cfg(None, optimize=1, fraction=0.25, maxnfunc=8, max_framelen=1, search=7), [raw], 44100)


def test_decoder_refuses_corrupt_headers_and_never_delivers_unverified_audio(engine):
    """untrusted input on the decode path: a bit plane count beyond the model's 32 planes, a frame capacity that overflows, an
    unknown arithmetic variant, a damaged payload -- every one is an error (no output), never out-of-bounds work or silently
    wrong audio (the audio MD5 decides, include/sac_b200.h)"""
    pcm = synth_pcm(1, 2, 61).astype(np.int32)[:6000]
    wav = _wav_bytes(pcm, sr=8000)
    sac, st = engine.encode_memory(sb.make_cfg("normal", max_framelen=1), wav)
    assert sac[17] == 1                                             # arithmetic variant: canonical
    back, st2 = engine.decode_memory(sac, len(wav) + 64)
    assert back == wav and st2.md5_ok == 1
    mds = struct.unpack("<I", sac[18:22])[0]
    first = 22 + mds + 16                                           # first frame record
    blk = first + 4 + 58 * 4                                        # first block header: u32 size, 3 x i32, u16 flag
    cases = {}
    b = bytearray(sac); b[blk + 16] = 40; cases["bit plane count out of range"] = bytes(b)
    b = bytearray(sac); b[blk + 16] = 31; cases["bit plane count out of range "] = bytes(b)
    b = bytearray(sac); b[16] = 255; b[6:10] = struct.pack("<I", 0x7fffffff); cases["frame length out of range"] = bytes(b)
    b = bytearray(sac); b[17] = 7; cases["unknown arithmetic variant"] = bytes(b)
    b = bytearray(sac); b[blk + 18 + 40] ^= 0x55; cases["MD5 mismatch"] = bytes(b)
    b = bytearray(sac); b[first:first + 4] = struct.pack("<I", 0x7fffffff); cases["bad sample count"] = bytes(b)
    for what, img in cases.items():
        with pytest.raises(sb.SacError, match=what.strip()):
            engine.decode_memory(img, len(wav) + 64)
    # a variant-0 file (what a reference build writes) is attempted: this short file decodes and verifies
    b = bytearray(sac); b[17] = 0
    back, st3 = engine.decode_memory(bytes(b), len(wav) + 64)
    assert back == wav and st3.md5_ok == 1
    # the engine is still usable after the refusals
    back, _ = engine.decode_memory(sac, len(wav) + 64)
    assert back == wav
