import hashlib

import numpy as np

import oracle_lib as ol
from synth_wav import synth_pcm


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def reference_view(img):
    """a .sac image (or its leading bytes) written by this library, as the reference would have written it: byte 17 of the
    header names the arithmetic variant (1 = canonical, include/sac_b200.h); the reference writes 0 there"""
    b = bytearray(img)
    assert b[:4] == b"SAC2" and b[17] == 1, "files of this library carry arithmetic variant 1 in header byte 17"
    b[17] = 0
    return bytes(b)


def case_profile(case, vmin, vmax, vdef):
    """same construction as tests/golden/make_golden.py"""
    if case["profile"] == "default":
        return vdef.copy()
    rng = np.random.default_rng(case["profile_seed"])
    u = rng.random(58).astype(np.float32)
    p = (vmin + u * (vmax - vmin)).astype(np.float32)
    p[28] = min(p[28], 2000); p[31] = min(p[31], 1500)
    if case.get("neg_ns1"):
        p[27] = -abs(p[27]) - 1
    return p


def case_planes(case):
    pcm = synth_pcm(case["secs"], case["nch"], case["seed"]).astype(np.int32)
    return ol.analyse([pcm[:, ch] for ch in range(case["nch"])])


def random_profile(rng, vmin, vmax, cap0=2000, cap1=1500):
    u = rng.random(58).astype(np.float32)
    p = (vmin + u * (vmax - vmin)).astype(np.float32)
    if cap0: p[28] = min(p[28], cap0)
    if cap1: p[31] = min(p[31], cap1)
    return p


def special_streams():
    rng = np.random.default_rng(99)
    return {
        "zeros": np.zeros(500, np.int32), "ones": np.ones(300, np.int32), "single": np.array([5], np.int32),
        "ramp": np.arange(2000, dtype=np.int32) % 97, "spikes": (rng.random(3000) < 0.02).astype(np.int32) * 30000,
        "laplace": np.abs(rng.laplace(0, 300, 5000)).astype(np.int32), "wide": rng.integers(0, 1 << 17, 1500).astype(np.int32),
    }


def oracle_file_image(wav_bytes, modes, sparse=1, optimize=None):
    """A whole .sac image the way the codec would write it: the host container plan (header, metadata, MD5, frame plan)
    followed by the CPU restatement's frame records; optimize = dict(fraction, maxnfunc, sigma, cost_kind[, num_threads])
    runs the DDS search per frame, each frame starting from the previous frame's optimum (libsac.cpp:461-476).
    Returns (image bytes, per-block mapped flags)."""
    import ctypes as C
    import wave, io
    import numpy as np
    import oracle_lib as ol
    import sac_b200 as sb
    lib = ol.oracle()
    lib.saco_set_modes(*modes)
    prefix, frames, st = sb.container_plan(sb.make_cfg("normal"), wav_bytes)
    with wave.open(io.BytesIO(wav_bytes)) as w:
        assert w.getsampwidth() == 2
        pcm = np.frombuffer(w.readframes(w.getnframes()), "<i2").astype(np.int32).reshape(-1, w.getnchannels())
    _, _, vdef = sb.base_profile()
    o = optimize or {}
    cfg = (C.c_int * 8)(1 if optimize else 0, int(round(o.get("fraction", 0) * 1e6)), o.get("maxnfunc", 0), o.get("num_threads", 0),
                        int(round(o.get("sigma", 0.2) * 1e6)), o.get("optk", 4), o.get("cost_kind", 2), 20 * st.samplerate)
    img, flags, pos = bytearray(prefix), [], 0
    prof = vdef.copy()
    for n in frames:
        s = [np.ascontiguousarray(pcm[pos:pos + n, ch]) for ch in range(st.nch)]
        out = np.zeros(8 * n + 4096, np.uint8)
        mapped = (C.c_int * 2)(0, 0)
        nb = lib.saco_encode_frame2(st.nch, n, ol._p(s[0], ol._i32p), ol._p(s[1], ol._i32p) if st.nch > 1 else None, ol._p(prof, ol._f32p), cfg,
                                    int(sparse), ol._p(out, ol._u8p), len(out), mapped)
        img += out[:nb].tobytes()
        flags += list(mapped)[:st.nch]
        pos += n
    return bytes(img), flags
