import hashlib

import numpy as np

import oracle_lib as ol
from synth_wav import synth_pcm


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


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
