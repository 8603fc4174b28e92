"""Generates tests/golden/golden.json from the REFERENCE ITSELF (oracle/_ref/libsacref_nc.so = the reference's own
classes compiled from /root/reference with -ffp-contract=off, see oracle/Makefile). Run in the build container only:

    make -C oracle ref && python tests/golden/make_golden.py

Each case records what the reference computes for a seeded synthetic input (BASELINE.md section 3 generator), so
that the restatement in oracle/sac_oracle.cpp can be pinned without /root/reference being present.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402
from synth_wav import synth_pcm  # noqa: E402


def canon_trace(trace):
    """population generations are evaluated by concurrent threads in the reference: hash the points in sorted order"""
    t = np.stack(trace)
    return t[np.lexsort(t.T[::-1])]


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def profile_for(case, vmin, vmax, vdef):
    if case["profile"] == "default":
        return vdef.copy()
    rng = np.random.default_rng(case["profile_seed"])
    u = rng.random(58).astype(np.float32)
    p = (vmin + u * (vmax - vmin)).astype(np.float32)
    p[28] = min(p[28], 2000); p[31] = min(p[31], 1500)
    if case.get("neg_ns1"):
        p[27] = -abs(p[27]) - 1
    return p


CASES = [
    dict(name="mono_default_k1", nch=1, secs=2, seed=1, profile="default", k=1, frm=0, n=30000),
    dict(name="mono_default_k4", nch=1, secs=2, seed=1, profile="default", k=4, frm=5000, n=30000),
    dict(name="stereo_default_k4", nch=2, secs=2, seed=2, profile="default", k=4, frm=1000, n=25000),
    dict(name="stereo_rand11_k4", nch=2, secs=2, seed=2, profile="random", profile_seed=11, k=4, frm=1000, n=20000),
    dict(name="stereo_rand12_k1", nch=2, secs=2, seed=3, profile="random", profile_seed=12, k=1, frm=0, n=12000),
    dict(name="stereo_rand13_swap_k4", nch=2, secs=2, seed=3, profile="random", profile_seed=13, neg_ns1=True, k=4, frm=300, n=20000),
    dict(name="stereo_tiny", nch=2, secs=1, seed=4, profile="default", k=4, frm=0, n=37),
]


def main():
    ref = ol.ref_lib(nc=True)
    assert ref is not None, "build oracle/_ref first (make -C oracle ref)"
    vmin, vmax, vdef = ol.base_profile(ref)
    out = {"generator": "tests/golden/make_golden.py", "reference": "slmdev/sac v0.7.25, -O3 -march=x86-64-v3 -ffp-contract=off",
           "base_profile": {"vmin": vmin.tolist(), "vmax": vmax.tolist(), "vdef": vdef.tolist()}, "predict": [], "bitplane": [],
           "laplace": [], "dds": [], "frame": []}
    for c in CASES:
        pcm = synth_pcm(c["secs"], c["nch"], c["seed"]).astype(np.int32)
        rf = ol.RefFrame(ref, c["nch"], 20 * 44100)
        rf.set_samples([pcm[:, ch] for ch in range(c["nch"])])
        rf.analyse()
        prof = profile_for(c, vmin, vmax, vdef)
        e = rf.predict_window(prof, c["frm"], c["n"], c["k"] > 1)
        assert c["k"] in (1, 4)
        rec = dict(c)
        rec["stats"] = [rf.stats(ch)[:3] for ch in range(c["nch"])]
        rec["fnv"] = [ol.fnv1a64(x) for x in e]
        rec["costs"] = {str(kind): [float(ref.ref_cost(kind, ol._p(x, ol._i32p), len(x))) for x in e] for kind in range(5)}
        out["predict"].append(rec)
        # bitplane payload of the first channel's residuals
        u = ol.s2u(e[0])
        maxbpn = ol.ilog2(int(u.max()))
        cap = 4 * len(u) + 64
        buf = np.zeros(cap, np.uint8)
        nb = ref.ref_bitplane_encode(ol._p(u, ol._i32p), len(u), maxbpn, ol._p(buf, ol._u8p), cap)
        out["bitplane"].append(dict(name=c["name"], maxbpn=maxbpn, nbytes=int(nb), sha1=sha(buf[:nb])))
    # pathological bitplane inputs
    rng = np.random.default_rng(99)
    specials = {
        "zeros": np.zeros(500, np.int32), "ones": np.ones(300, np.int32), "single": np.array([5], np.int32),
        "ramp": np.arange(2000, dtype=np.int32) % 97, "spikes": (rng.random(3000) < 0.02).astype(np.int32) * 30000,
        "laplace": np.abs(rng.laplace(0, 300, 5000)).astype(np.int32), "wide": rng.integers(0, 1 << 17, 1500).astype(np.int32),
    }
    for name, u in specials.items():
        maxbpn = ol.ilog2(int(u.max()))
        cap = 4 * len(u) + 64
        buf = np.zeros(cap, np.uint8)
        nb = ref.ref_bitplane_encode(ol._p(u, ol._i32p), len(u), maxbpn, ol._p(buf, ol._u8p), cap)
        out["bitplane"].append(dict(name="special_" + name, maxbpn=maxbpn, nbytes=int(nb), sha1=sha(buf[:nb]), seed=99))
    # PredictLaplace samples
    for bpn in (0, 1, 5, 9, 13, 16):
        for avg in (0, 1, 2, 3, 17, 255, 1000, 4097, 65535, 131071):
            out["laplace"].append([avg, bpn, int(ref.ref_predict_laplace(avg, bpn))])
    # DDS on a synthetic objective: the evaluated points and the result pin the <random> call sequence
    D = 56
    xmin = np.array([vmin[i] for i in range(58) if i not in (56, 57)], np.float64)
    xmax = np.array([vmax[i] for i in range(58) if i not in (56, 57)], np.float64)
    xs = np.array([vdef[i] for i in range(58) if i not in (56, 57)], np.float64)
    for nfunc, nt, sigma in ((60, 0, 0.2), (200, 0, 0.25), (120, 6, 0.2), (100, 16, 0.25)):
        trace = []

        def cb(xp, n, _u):
            x = np.ctypeslib.as_array(xp, shape=(n,)).copy()
            trace.append(x)
            z = (x - xmin) / (xmax - xmin)
            return float(np.sum((z - 0.37) ** 2) + 0.05 * np.sum(np.cos(9 * z)))

        fn = ol.COST_CB(cb)
        xb = np.zeros(D)
        fb = ref.ref_dds_run(D, ol._p(xmin, ol._f64p), ol._p(xmax, ol._f64p), ol._p(xs, ol._f64p), nfunc, nt, sigma, fn, None,
                             ol._p(xb, ol._f64p))
        out["dds"].append(dict(nfunc=nfunc, num_threads=nt, sigma=sigma, best=float(fb), xbest_sha1=sha(xb),
                               trace_sha1=sha(canon_trace(trace)), evals=len(trace)))
    # whole-frame records (Predict + Encode) without and with a short search
    for nm, nch, secs, seed, kw in (("frame_mono_normal", 1, 1, 21, dict(optimize=0)),
                                   ("frame_stereo_normal", 2, 1, 22, dict(optimize=0)),
                                   ("frame_stereo_dds8", 2, 1, 23, dict(optimize=1, fraction=0.02, maxnfunc=8, sigma=0.2, cost_kind=2)),
                                   ("frame_stereo_dds6_bpn_pop3", 2, 1, 24, dict(optimize=1, fraction=0.015, maxnfunc=7, num_threads=3, sigma=0.25, cost_kind=4))):
        pcm = synth_pcm(secs, nch, seed).astype(np.int32)
        rf = ol.RefFrame(ref, nch, 20 * 44100, **kw)
        rf.set_samples([pcm[:, ch] for ch in range(nch)])
        rf.predict(); rf.encode()
        rec = dict(name=nm, nch=nch, secs=secs, seed=seed, cfg=kw, profile_sha1=sha(rf.profile()),
                   stats=[rf.stats(ch) for ch in range(nch)], payload_sha1=[sha(rf.encoded(ch)) for ch in range(nch)],
                   payload_len=[int(len(rf.encoded(ch))) for ch in range(nch)])
        out["frame"].append(rec)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote golden.json:", {k: len(v) for k, v in out.items() if isinstance(v, list)})


if __name__ == "__main__":
    main()
