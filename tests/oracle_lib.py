"""ctypes access to the checkers: oracle/liboracle.so (our CPU restatement) and, when present,
oracle/_ref/libsacref*.so (the reference's own classes compiled from /root/reference).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
sys.path.insert(0, os.path.join(ROOT, "tools"))

ORDER_REF, ORDER_B200 = 0, 1
MATH_LIBM, MATH_CANON = 0, 1
COST_L1, COST_RMS, COST_ENTROPY, COST_GOLOMB, COST_BITPLANE = range(5)

_i32p = C.POINTER(C.c_int32)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_u16p = C.POINTER(C.c_uint16)
COST_CB = C.CFUNCTYPE(C.c_double, _f64p, C.c_int, C.c_void_p)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "liboracle.so"])


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        lib = C.CDLL(path)
        lib.saco_cost.restype = C.c_double
        lib.saco_dds_run.restype = C.c_double
        lib.saco_dds_run.argtypes = [C.c_int, _f64p, _f64p, _f64p, C.c_int, C.c_int, C.c_double, COST_CB, C.c_void_p, _f64p]
        _oracle = lib
    return _oracle


def ref_lib(nc=True):
    """reference harness; nc=True -> the -ffp-contract=off build. None if not built (no /root/reference, no prebuilt)."""
    path = os.path.join(ORACLE_DIR, "_ref", "libsacref_nc.so" if nc else "libsacref.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    lib.ref_frame_new.restype = C.c_void_p
    lib.ref_frame_new.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.c_int]
    lib.ref_cost.restype = C.c_double
    lib.ref_dds_run.restype = C.c_double
    lib.ref_dds_run.argtypes = [C.c_int, _f64p, _f64p, _f64p, C.c_int, C.c_int, C.c_double, COST_CB, C.c_void_p, _f64p]
    if hasattr(lib, "ref_cma_run"):
        lib.ref_cma_run.restype = C.c_double
        lib.ref_cma_run.argtypes = [C.c_int, _f64p, _f64p, _f64p, C.c_int, C.c_double, COST_CB, C.c_void_p, _f64p]
    if hasattr(lib, "ref_de_run"):
        lib.ref_de_run.restype = C.c_double
        lib.ref_de_run.argtypes = [C.c_int, _f64p, _f64p, _f64p, C.c_int, C.c_double, COST_CB, C.c_void_p, _f64p]
    for f in ("ref_frame_free", "ref_frame_set_samples", "ref_frame_analyse", "ref_frame_get_stats", "ref_frame_set_stats",
              "ref_frame_set_profile", "ref_frame_get_profile", "ref_frame_predict_window", "ref_frame_predict",
              "ref_frame_encode", "ref_frame_get_error", "ref_frame_get_encoded", "ref_frame_set_mt"):
        getattr(lib, f).argtypes = None
    if hasattr(lib, "ref_frame_load_block"):
        lib.ref_frame_load_block.argtypes = [C.c_void_p, C.c_int, _u8p, C.c_int, C.c_int32, C.c_int32, C.c_int32, C.c_int, C.c_int]
        lib.ref_frame_decode.argtypes = [C.c_void_p, C.c_int]
        lib.ref_frame_get_samples.argtypes = [C.c_void_p, C.c_int, _i32p]
        lib.ref_frame_get_maxbpn_map.argtypes = [C.c_void_p, C.c_int]
    return lib


def ref_sac_binary():
    p = os.path.join(ORACLE_DIR, "_ref", "sac")
    return p if os.path.exists(p) else None


def base_profile(lib=None):
    lib = lib or oracle()
    vmin = np.zeros(58, np.float32); vmax = np.zeros(58, np.float32); vdef = np.zeros(58, np.float32)
    fn = lib.saco_base_profile if hasattr(lib, "saco_base_profile") else lib.ref_base_profile
    fn(_p(vmin, _f32p), _p(vmax, _f32p), _p(vdef, _f32p))
    return vmin, vmax, vdef


def fnv1a64(e):
    """h ^= (uint32)e; h *= 1099511628211 over int32 residuals."""
    h = 1469598103934665603
    for v in np.asarray(e).astype(np.uint32).tolist():
        h = ((h ^ v) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


def analyse(planes):
    """mean / zero-mean / min / max per channel as FrameCoder::Predict's prologue (libsac.cpp:443-459).
    planes: list of int32 arrays. Returns (mean-free planes, means, minmax[2*nch])."""
    out, means, mm = [], [], []
    for s in planes:
        s = np.asarray(s, np.int32)
        n = len(s)
        mean = int(np.floor(int(s.astype(np.int64).sum()) / float(n))) if n else 0
        z = (s - mean).astype(np.int32)
        out.append(np.ascontiguousarray(z)); means.append(mean); mm += [int(z.min()), int(z.max())]
    return out, means, np.array(mm, np.int32)


def oracle_predict(planes, mm, profile, k, frm=0, n=None, order=ORDER_B200, math=MATH_CANON):
    lib = oracle()
    lib.saco_set_modes(order, math)
    nch = len(planes)
    n = len(planes[0]) - frm if n is None else n
    e = [np.zeros(n, np.int32) for _ in range(nch)]
    prof = np.ascontiguousarray(profile, np.float32)
    mm = np.ascontiguousarray(mm, np.int32)
    rc = lib.saco_predict_frame(nch, _p(planes[0], _i32p), _p(planes[1], _i32p) if nch > 1 else None, frm, n, _p(prof, _f32p),
                                k, _p(mm, _i32p), _p(e[0], _i32p), _p(e[1], _i32p) if nch > 1 else None)
    return e, rc


def oracle_unpredict(errs, mm, profile, order=ORDER_B200, math=MATH_CANON):
    lib = oracle()
    lib.saco_set_modes(order, math)
    nch = len(errs); n = len(errs[0])
    s = [np.zeros(n, np.int32) for _ in range(nch)]
    prof = np.ascontiguousarray(profile, np.float32)
    mm = np.ascontiguousarray(mm, np.int32)
    lib.saco_unpredict_frame(nch, n, _p(prof, _f32p), _p(mm, _i32p), _p(errs[0], _i32p), _p(errs[1], _i32p) if nch > 1 else None,
                             _p(s[0], _i32p), _p(s[1], _i32p) if nch > 1 else None)
    return s


def oracle_cost(kind, e, math=MATH_CANON):
    lib = oracle()
    lib.saco_set_modes(ORDER_B200, math)
    e = np.ascontiguousarray(e, np.int32)
    return lib.saco_cost(kind, _p(e, _i32p), len(e))


def s2u(e):
    e = np.asarray(e, np.int64)
    return np.where(e < 0, -2 * e, np.where(e > 0, 2 * e - 1, 0)).astype(np.int32)


def ilog2(v):
    return int(v).bit_length() - 1 if v > 0 else 0


def oracle_bitplane_encode(u, maxbpn=None, math=MATH_CANON):
    lib = oracle()
    lib.saco_set_modes(ORDER_B200, math)
    u = np.ascontiguousarray(u, np.int32)
    if maxbpn is None:
        maxbpn = ilog2(int(u.max()) if len(u) else 0)
    cap = 4 * len(u) + 64
    out = np.zeros(cap, np.uint8)
    nb = lib.saco_bitplane_encode(_p(u, _i32p), len(u), maxbpn, _p(out, _u8p), cap)
    return out[:nb].copy(), maxbpn


def oracle_bitplane_decode(payload, n, maxbpn, math=MATH_CANON):
    lib = oracle()
    lib.saco_set_modes(ORDER_B200, math)
    payload = np.ascontiguousarray(payload, np.uint8)
    out = np.zeros(n, np.int32)
    lib.saco_bitplane_decode(_p(payload, _u8p), len(payload), n, maxbpn, _p(out, _i32p))
    return out


class RefFrame:
    """thin wrapper over the reference FrameCoder through the harness"""

    def __init__(self, lib, nch, framesize, optimize=0, fraction=0.0, maxnfunc=0, num_threads=0, sigma=0.2, optk=4,
                 cost_kind=COST_ENTROPY, reset=0, sparse_pcm=1, zero_mean=1):
        self.lib, self.nch = lib, nch
        # FrameCoder::SearchCost enum order is L1,RMS,Entropy,Golomb,Bitplane == ours
        self.h = C.c_void_p(lib.ref_frame_new(nch, framesize, optimize, fraction, maxnfunc, num_threads, sigma, optk,
                                              cost_kind, reset, sparse_pcm, zero_mean))
        self.n = 0

    def set_samples(self, planes):
        for ch, s in enumerate(planes):
            s = np.ascontiguousarray(s, np.int32)
            self.lib.ref_frame_set_samples(self.h, ch, _p(s, _i32p), len(s))
            self.n = len(s)

    def analyse(self):
        self.lib.ref_frame_analyse(self.h)

    def stats(self, ch):
        st = (C.c_int32 * 6)()
        self.lib.ref_frame_get_stats(self.h, ch, st)
        return list(st)

    def predict_window(self, profile, frm, n, optimize):
        prof = np.ascontiguousarray(profile, np.float32)
        e = [np.zeros(n, np.int32) for _ in range(self.nch)]
        self.lib.ref_frame_predict_window(self.h, _p(prof, _f32p), frm, n, int(optimize), _p(e[0], _i32p),
                                          _p(e[1], _i32p) if self.nch > 1 else None)
        return e

    def predict(self):
        self.lib.ref_frame_predict(self.h)

    def encode(self):
        self.lib.ref_frame_encode(self.h)

    def profile(self):
        v = np.zeros(58, np.float32)
        self.lib.ref_frame_get_profile(self.h, _p(v, _f32p))
        return v

    def error(self, ch):
        e = np.zeros(self.n, np.int32)
        self.lib.ref_frame_get_error(self.h, ch, _p(e, _i32p))
        return e

    def encoded(self, ch):
        nb = self.lib.ref_frame_get_encoded(self.h, ch, None, 0)
        out = np.zeros(nb, np.uint8)
        self.lib.ref_frame_get_encoded(self.h, ch, _p(out, _u8p), nb)
        return out

    def maxbpn_map(self, ch):
        return int(self.lib.ref_frame_get_maxbpn_map(self.h, ch))

    def decode_blocks(self, n, profile, blocks):
        """blocks: per channel (payload u8, mean, min, max, maxbpn, mapped) -> the reference's Decode()+Unpredict() samples"""
        prof = np.ascontiguousarray(profile, np.float32)
        self.lib.ref_frame_set_profile(self.h, _p(prof, _f32p))
        for ch, (pay, mean, mn, mx, bpn, mapped) in enumerate(blocks):
            pay = np.ascontiguousarray(pay, np.uint8)
            self.lib.ref_frame_load_block(self.h, ch, _p(pay, _u8p), len(pay), int(mean), int(mn), int(mx), int(bpn), int(mapped))
        self.lib.ref_frame_decode(self.h, n)
        out = []
        for ch in range(self.nch):
            d = np.zeros(n, np.int32)
            self.lib.ref_frame_get_samples(self.h, ch, _p(d, _i32p))
            out.append(d)
        return out

    def __del__(self):
        try:
            self.lib.ref_frame_free(self.h)
        except Exception:
            pass
