"""sac_b200 -- Python binding (ctypes) of libsac_b200.so, the B200-native encode path of the Sac lossless audio codec.

The product is the C ABI in include/sac_b200.h (CUDA kernels for sm_100a behind plain pointers and sizes); this
module only loads it and mirrors the reference's frame-coder surface (/root/reference src/libsac/libsac.h:45-56,
src/opt/opt.h:17-25) for tests and bench.py. There is no CPU fallback: without the built library or without a GPU
every compute call raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SAC_B200_LIB") or os.path.join(_HERE, "libsac_b200.so")   # the override selects a build variant (tools/ probes)
CLI_PATH = os.path.join(_HERE, "sac")

PROFILE_SIZE = 58
SEARCH_DIMS = [i for i in range(58) if i not in (56, 57)]
COST_L1, COST_RMS, COST_ENTROPY, COST_GOLOMB, COST_BITPLANE = range(5)
SEARCH_DDS, SEARCH_DE, SEARCH_CMA = range(3)

_i32p = C.POINTER(C.c_int32)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)
_intp = C.POINTER(C.c_int)
EVAL_FN = C.CFUNCTYPE(C.c_int, _f64p, C.c_int, C.c_int, _f64p, C.c_void_p)


class SacError(RuntimeError):
    pass


class Cfg(C.Structure):
    """sac_cfg (FrameCoder::tsac_cfg / toptim_cfg, libsac.h:19-44)"""
    _fields_ = [("optimize", C.c_int), ("fraction", C.c_double), ("maxnfunc", C.c_int), ("num_threads", C.c_int),
                ("sigma", C.c_double), ("optk", C.c_int), ("cost_kind", C.c_int), ("reset", C.c_int), ("zero_mean", C.c_int),
                ("sparse_pcm", C.c_int), ("max_framelen", C.c_int), ("adapt_block", C.c_int), ("frame_parallel", C.c_int),
                ("verbose", C.c_int), ("search", C.c_int), ("spec", C.c_int), ("inflight", C.c_int), ("grade", C.c_int)]


class FileStats(C.Structure):
    _fields_ = [("in_bytes", C.c_longlong), ("out_bytes", C.c_longlong), ("numsamples", C.c_int), ("nch", C.c_int),
                ("samplerate", C.c_int), ("bits", C.c_int), ("nframes", C.c_int), ("seconds", C.c_double),
                ("md5", C.c_uint8 * 16), ("md5_ok", C.c_int)]


def build(verbose=False):
    """compile libsac_b200.so and the sac CLI in-tree for sm_100a (nvcc cross-compiles without a GPU)"""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"] + ([] if verbose else ["-s"])
    subprocess.check_call(cmd)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SacError("libsac_b200.so is not built (run sac_b200.build() / make -C sac_b200/csrc); there is no fallback path")
        L = C.CDLL(LIB_PATH)
        L.sac_version.restype = C.c_char_p
        L.sac_last_error.restype = C.c_char_p
        L.sac_engine_create.restype = C.c_void_p
        L.sac_engine_create.argtypes = [C.c_int]
        L.sac_engine_destroy.argtypes = [C.c_void_p]
        L.sac_engine_launches.restype = C.c_longlong
        L.sac_engine_launches.argtypes = [C.c_void_p]
        L.sac_engine_last_timing.argtypes = [C.c_void_p, _f64p, C.POINTER(C.c_longlong)]
        L.sac_engine_total_timing.argtypes = [C.c_void_p, _f64p, C.POINTER(C.c_longlong)]
        L.sac_engine_set_dedup.argtypes = [C.c_void_p, C.c_int]
        L.sac_engine_set_grade.argtypes = [C.c_void_p, C.c_int]
        L.sac_engine_grade_stats.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
        L.sac_dedup_totals.argtypes = [C.POINTER(C.c_longlong)]
        L.sac_window_create.restype = C.c_void_p
        L.sac_window_create.argtypes = [C.c_void_p, C.c_int, C.POINTER(_i32p), C.c_int, _i32p]
        L.sac_window_destroy.argtypes = [C.c_void_p]
        L.sac_predict.argtypes = [C.c_void_p, C.c_void_p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, _intp]
        L.sac_cost.argtypes = [C.c_void_p, C.c_int, _i32p, C.c_int, C.c_int, _f64p]
        L.sac_bitplane_encode.argtypes = [C.c_void_p, _i32p, C.c_int, _intp, _u8p, C.c_longlong, C.POINTER(C.c_longlong)]
        L.sac_bitplane_decode.argtypes = [C.c_void_p, _u8p, C.c_longlong, C.c_int, C.c_int, _i32p]
        L.sac_eval_population.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, _f32p, _intp, C.c_int, _f64p, C.c_int,
                                          C.c_int, C.c_int, _f64p]
        L.sac_eval_jobs.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), _intp, _intp, _f32p, _intp, C.c_int, _f64p,
                                    C.c_int, C.c_int, _f64p]
        L.sac_dds_run.restype = C.c_double
        L.sac_dds_run.argtypes = [C.c_int, _f64p, _f64p, _f64p, C.c_int, C.c_int, C.c_double, EVAL_FN, C.c_void_p, _f64p]
        L.sac_dds_run_spec.restype = C.c_double
        L.sac_dds_run_spec.argtypes = [C.c_int, _f64p, _f64p, _f64p, C.c_int, C.c_double, C.c_int, EVAL_FN, C.c_void_p, _f64p,
                                       C.POINTER(C.c_longlong)]
        L.sac_analyse_subframes.argtypes = [C.c_int, C.POINTER(_i32p), C.c_int, C.c_int, _intp, _intp, _intp, C.c_int]
        L.sac_de_run.restype = C.c_double
        L.sac_de_run.argtypes = [C.c_int, _f64p, _f64p, _f64p, C.c_int, C.c_double, EVAL_FN, C.c_void_p, _f64p]
        L.sac_cma_run.restype = C.c_double
        L.sac_cma_run.argtypes = [C.c_int, _f64p, _f64p, _f64p, C.c_int, C.c_double, EVAL_FN, C.c_void_p, _f64p]
        L.sac_cfg_default.argtypes = [C.POINTER(Cfg)]
        L.sac_cfg_preset.argtypes = [C.POINTER(Cfg), C.c_char_p]
        L.sac_frames_encode.argtypes = [C.c_void_p, C.POINTER(Cfg), C.c_int, C.c_int, C.c_int, C.POINTER(_i32p), _intp, _f32p,
                                        _u8p, C.c_longlong, C.POINTER(C.c_longlong)]
        L.sac_frames_encode_resident.argtypes = [C.c_void_p, C.POINTER(Cfg), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), _i32p,
                                                 _f32p, _u8p, C.c_longlong, C.POINTER(C.c_longlong)]
        L.sac_fp64_peak_gflops.restype = C.c_double
        L.sac_fp64_peak_gflops.argtypes = [C.c_void_p]
        L.sac_frame_decode.restype = C.c_longlong
        L.sac_frame_decode.argtypes = [C.c_void_p, C.c_int, _u8p, C.c_longlong, C.POINTER(_i32p), C.c_int, _intp]
        L.sac_encode_file.argtypes = [C.c_void_p, C.POINTER(Cfg), C.c_char_p, C.c_char_p, C.POINTER(FileStats)]
        L.sac_decode_file.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(FileStats)]
        L.sac_encode_file_multi.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(Cfg), C.c_char_p, C.c_char_p, C.POINTER(FileStats)]
        L.sac_encode_files.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(Cfg), C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p),
                                       C.POINTER(FileStats), _intp]
        L.sac_encode_memory.argtypes = [C.c_void_p, C.POINTER(Cfg), _u8p, C.c_longlong, _u8p, C.c_longlong,
                                        C.POINTER(C.c_longlong), C.POINTER(FileStats)]
        L.sac_decode_memory.argtypes = [C.c_void_p, _u8p, C.c_longlong, _u8p, C.c_longlong, C.POINTER(C.c_longlong),
                                        C.POINTER(FileStats)]
        _lib = L
    return _lib


def _chk(rc, what):
    if rc != 0:
        raise SacError("%s failed (%d): %s" % (what, rc, lib().sac_last_error().decode()))


def _p(a, t):
    return a.ctypes.data_as(t)


def base_profile():
    vmin = np.zeros(58, np.float32); vmax = np.zeros(58, np.float32); vdef = np.zeros(58, np.float32)
    lib().sac_base_profile(_p(vmin, _f32p), _p(vmax, _f32p), _p(vdef, _f32p))
    return vmin, vmax, vdef


def make_cfg(preset=None, **kw):
    c = Cfg()
    lib().sac_cfg_default(C.byref(c))
    if preset:
        _chk(lib().sac_cfg_preset(C.byref(c), preset.encode()), "sac_cfg_preset")
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def dds_run(func, xmin, xmax, xstart, nfunc_max, num_threads=0, sigma_init=0.2):
    """OptDDS::run with a Python population evaluator func(X[P,D]) -> costs[P] (host only, no GPU needed)"""
    xmin = np.ascontiguousarray(xmin, np.float64); xmax = np.ascontiguousarray(xmax, np.float64)
    xstart = np.ascontiguousarray(xstart, np.float64)
    D = len(xstart)
    xbest = np.zeros(D)

    def cb(Xp, P, Dd, costp, _user):
        X = np.ctypeslib.as_array(Xp, shape=(P, Dd))
        out = np.ctypeslib.as_array(costp, shape=(P,))
        out[:] = np.asarray(func(X.copy()), np.float64)
        return 0

    fn = EVAL_FN(cb)
    best = lib().sac_dds_run(D, _p(xmin, _f64p), _p(xmax, _f64p), _p(xstart, _f64p), nfunc_max, num_threads, sigma_init, fn, None,
                             _p(xbest, _f64p))
    return best, xbest


def dds_run_spec(func, xmin, xmax, xstart, nfunc_max, sigma_init=0.2, spec=16):
    """run_single in speculative batches (host only): returns (best cost, best x, candidates evaluated)"""
    xmin = np.ascontiguousarray(xmin, np.float64); xmax = np.ascontiguousarray(xmax, np.float64)
    xstart = np.ascontiguousarray(xstart, np.float64)
    D = len(xstart)
    xbest = np.zeros(D)

    def cb(Xp, P, Dd, costp, _user):
        X = np.ctypeslib.as_array(Xp, shape=(P, Dd))
        out = np.ctypeslib.as_array(costp, shape=(P,))
        out[:] = np.asarray(func(X.copy()), np.float64)
        return 0

    fn = EVAL_FN(cb)
    ev = C.c_longlong(0)
    best = lib().sac_dds_run_spec(D, _p(xmin, _f64p), _p(xmax, _f64p), _p(xstart, _f64p), nfunc_max, sigma_init, spec, fn, None,
                                  _p(xbest, _f64p), C.byref(ev))
    return best, xbest, int(ev.value)


def encode_file_multi(engines, cfg, wav_path, sac_path):
    """one file on several GPUs (one Engine each): frames dealt round-robin, --opt-reset semantics"""
    arr = (C.c_void_p * len(engines))(*[e.h for e in engines])
    st = FileStats()
    _chk(lib().sac_encode_file_multi(arr, len(engines), C.byref(cfg), str(wav_path).encode(), str(sac_path).encode(), C.byref(st)), "sac_encode_file_multi")
    return st


def encode_files(engines, cfg, wav_paths, sac_paths):
    """a batch of files over several GPUs: returns (per-file FileStats, per-file status)"""
    n = len(wav_paths)
    arr = (C.c_void_p * len(engines))(*[e.h for e in engines])
    wp = (C.c_char_p * n)(*[str(p).encode() for p in wav_paths]); sp = (C.c_char_p * n)(*[str(p).encode() for p in sac_paths])
    st = (FileStats * max(n, 1))(); status = np.zeros(max(n, 1), np.int32)
    rc = lib().sac_encode_files(arr, len(engines), C.byref(cfg), n, wp, sp, st, _p(status, _intp))
    if rc:
        raise SacError("sac_encode_files failed (%d): %s" % (rc, lib().sac_last_error().decode()))
    return list(st)[:n], status[:n].tolist()


def analyse_subframes(planes, samplerate):
    """Codec::Analyse for one read (host only): [(start, length, state)]"""
    planes = [np.ascontiguousarray(p, np.int32) for p in planes]
    arr = (_i32p * len(planes))(*[_p(p, _i32p) for p in planes])
    cap = 64
    st = np.zeros(cap, np.int32); ln = np.zeros(cap, np.int32); ss = np.zeros(cap, np.int32)
    n = lib().sac_analyse_subframes(len(planes), arr, len(planes[0]), samplerate, _p(st, _intp), _p(ln, _intp), _p(ss, _intp), cap)
    if n < 0:
        raise SacError("sac_analyse_subframes failed (%d): %s" % (n, lib().sac_last_error().decode()))
    return [(int(st[i]), int(ln[i]), int(ss[i])) for i in range(n)]


def de_run(func, xmin, xmax, xstart, nfunc_max, sigma_init=0.15, entry="sac_de_run"):
    """OptDE::run (or, entry="sac_cma_run", OptCMA::run) with a Python population evaluator func(X[P,D]) -> costs[P]
    (host only, no GPU needed)"""
    xmin = np.ascontiguousarray(xmin, np.float64); xmax = np.ascontiguousarray(xmax, np.float64)
    xstart = np.ascontiguousarray(xstart, np.float64)
    D = len(xstart)
    xbest = np.zeros(D)

    def cb(Xp, P, Dd, costp, _user):
        X = np.ctypeslib.as_array(Xp, shape=(P, Dd))
        out = np.ctypeslib.as_array(costp, shape=(P,))
        out[:] = np.asarray(func(X.copy()), np.float64)
        return 0

    fn = EVAL_FN(cb)
    best = getattr(lib(), entry)(D, _p(xmin, _f64p), _p(xmax, _f64p), _p(xstart, _f64p), nfunc_max, sigma_init, fn, None, _p(xbest, _f64p))
    return best, xbest


def cma_run(func, xmin, xmax, xstart, nfunc_max, sigma_init=0.0):
    return de_run(func, xmin, xmax, xstart, nfunc_max, sigma_init, entry="sac_cma_run")


def profile_params(profile58):
    """host-only: FrameCoder::SetParam as this library maps a profile (59 doubles, order in include/sac_b200.h)"""
    prof = np.ascontiguousarray(profile58, np.float32)
    out = np.zeros(64, np.float64)
    n = lib().sac_profile_params(_p(prof, _f32p), _p(out, _f64p), 64)
    if n < 0:
        raise SacError("sac_profile_params failed: %s" % lib().sac_last_error().decode())
    return out[:n]


def frame_stats(samples, zero_mean=1):
    """host-only: (mean, min, max) of one channel as the block header carries them"""
    s = np.ascontiguousarray(samples, np.int32)
    out = np.zeros(3, np.int32)
    _chk(lib().sac_frame_stats(_p(s, _i32p), len(s), int(zero_mean), _p(out, _i32p)), "sac_frame_stats")
    return [int(x) for x in out]


def container_plan(cfg, wav_bytes, cap_frames=4096):
    """host-only: (.sac bytes preceding the first frame record, [samples per frame record], FileStats)"""
    wav = np.frombuffer(wav_bytes, np.uint8)
    cap = len(wav) + 65536
    out = np.zeros(cap, np.uint8)
    olen = C.c_longlong(0)
    fl = np.zeros(cap_frames, np.int32)
    st = FileStats()
    _chk(lib().sac_container_plan(C.byref(cfg), _p(wav, _u8p), len(wav), _p(out, _u8p), cap, C.byref(olen), _p(fl, _i32p), cap_frames,
                                  C.byref(st)), "sac_container_plan")
    return out[:olen.value].tobytes(), fl[:st.nframes].tolist(), st


class Window:
    """a (sub)frame resident in HBM: mean-free planes + stats (what FrameCoder::Optimize's cost lambda captures)"""

    def __init__(self, engine, planes, minmax):
        self.engine = engine
        self.nch = len(planes)
        self.n = len(planes[0])
        self._planes = [np.ascontiguousarray(p, np.int32) for p in planes]
        arr = (_i32p * self.nch)(*[_p(p, _i32p) for p in self._planes])
        mm = np.ascontiguousarray(minmax, np.int32)
        self.h = lib().sac_window_create(engine.h, self.nch, arr, self.n, _p(mm, _i32p))
        if not self.h:
            raise SacError("sac_window_create: " + lib().sac_last_error().decode())

    def close(self):
        if self.h:
            lib().sac_window_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    def __init__(self, device=0):
        self.h = lib().sac_engine_create(device)
        if not self.h:
            raise SacError("sac_engine_create: " + lib().sac_last_error().decode())

    def close(self):
        if self.h:
            lib().sac_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self):
        return int(lib().sac_engine_launches(self.h))

    def last_timing(self):
        ms = (C.c_double * 4)(); ln = (C.c_longlong * 4)()
        lib().sac_engine_last_timing(self.h, ms, ln)
        return list(ms), list(ln)

    def window(self, planes, minmax):
        return Window(self, planes, minmax)

    def predict(self, win, profiles, frm, n, k):
        """FrameCoder::PredictFrame for P profiles: returns residuals [P, nch, n] and flags [P]"""
        profiles = np.ascontiguousarray(np.atleast_2d(profiles), np.float32)
        P = profiles.shape[0]
        out = np.zeros((P, win.nch, n), np.int32)
        flags = np.zeros(P, np.int32)
        _chk(lib().sac_predict(self.h, win.h, _p(profiles, _f32p), P, frm, n, k, _p(out, _i32p), _p(flags, _intp)), "sac_predict")
        return out, flags

    def cost(self, kind, bufs):
        bufs = np.ascontiguousarray(np.atleast_2d(bufs), np.int32)
        out = np.zeros(bufs.shape[0])
        _chk(lib().sac_cost(self.h, kind, _p(bufs, _i32p), bufs.shape[0], bufs.shape[1], _p(out, _f64p)), "sac_cost")
        return out

    def bitplane_encode(self, resid, maxbpn=-1):
        """BitplaneCoder::Encode over RangeCoderSH: signed residuals -> (payload bytes, maxbpn)"""
        resid = np.ascontiguousarray(resid, np.int32)
        cap = 4 * len(resid) + 1024
        out = np.zeros(cap, np.uint8)
        olen = C.c_longlong(0)
        mb = C.c_int(maxbpn)
        _chk(lib().sac_bitplane_encode(self.h, _p(resid, _i32p), len(resid), C.byref(mb), _p(out, _u8p), cap, C.byref(olen)),
             "sac_bitplane_encode")
        return out[:olen.value].copy(), mb.value

    def bitplane_decode(self, payload, n, maxbpn):
        payload = np.ascontiguousarray(payload, np.uint8)
        out = np.zeros(n, np.int32)
        _chk(lib().sac_bitplane_decode(self.h, _p(payload, _u8p), len(payload), n, maxbpn, _p(out, _i32p)), "sac_bitplane_decode")
        return out

    def eval_population(self, win, frm, n, base, X, cost_kind, optk=4, dims=None):
        """Opt::eval_points_mt over the cost lambda of FrameCoder::Optimize: X[P,D] -> cost[P]"""
        X = np.ascontiguousarray(np.atleast_2d(X), np.float64)
        dims = np.ascontiguousarray(SEARCH_DIMS if dims is None else dims, np.int32)
        base = np.ascontiguousarray(base, np.float32)
        out = np.zeros(X.shape[0])
        _chk(lib().sac_eval_population(self.h, win.h, frm, n, _p(base, _f32p), _p(dims, _intp), X.shape[1], _p(X, _f64p), X.shape[0],
                                       cost_kind, optk, _p(out, _f64p)), "sac_eval_population")
        return out

    def frames_encode(self, cfg, frames, max_framesize, profile=None):
        """frames: list of lists of int32 planes (raw samples). Returns (bytes, profile_out)."""
        nch = len(frames[0])
        keep = [[np.ascontiguousarray(p, np.int32) for p in fr] for fr in frames]
        flat = [p for fr in keep for p in fr]
        arr = (_i32p * len(flat))(*[_p(p, _i32p) for p in flat])
        ns = np.array([len(fr[0]) for fr in keep], np.int32)
        prof = np.ascontiguousarray(base_profile()[2] if profile is None else profile, np.float32).copy()
        cap = int(sum(int(x) for x in ns) * nch * 5 + 4096 * len(keep))
        out = np.zeros(cap, np.uint8)
        olen = C.c_longlong(0)
        _chk(lib().sac_frames_encode(self.h, C.byref(cfg), nch, max_framesize, len(keep), arr, _p(ns, _intp), _p(prof, _f32p),
                                     _p(out, _u8p), cap, C.byref(olen)), "sac_frames_encode")
        return out[:olen.value].copy(), prof

    def frames_encode_resident(self, cfg, windows, means, max_framesize, profile=None):
        """frames already in HBM (Window objects of mean-free planes); means: per frame per channel"""
        nch = windows[0].nch
        arr = (C.c_void_p * len(windows))(*[w.h for w in windows])
        mm = np.ascontiguousarray(means, np.int32).reshape(-1)
        prof = np.ascontiguousarray(base_profile()[2] if profile is None else profile, np.float32).copy()
        cap = int(sum(w.n for w in windows) * nch * 5 + 4096 * len(windows))
        out = np.zeros(cap, np.uint8)
        olen = C.c_longlong(0)
        _chk(lib().sac_frames_encode_resident(self.h, C.byref(cfg), nch, max_framesize, len(windows), arr, _p(mm, _i32p), _p(prof, _f32p),
                                              _p(out, _u8p), cap, C.byref(olen)), "sac_frames_encode_resident")
        return out[:olen.value].copy(), prof

    def set_dedup(self, on):
        """exact de-duplication of identical chains / OLS stages within a call; returns the previous setting"""
        return int(lib().sac_engine_set_dedup(self.h, int(bool(on))))

    def total_timing(self):
        """device ms since creation by kernel class (ols, cascade, bitplane, other) over all streams + evaluations timed"""
        ms = (C.c_double * 4)(); calls = C.c_longlong(0)
        lib().sac_engine_total_timing(self.h, ms, C.byref(calls))
        return list(ms), int(calls.value)

    def set_grade(self, grade):
        """0 = canonical arithmetic, 1 = search-grade kernels for predict / eval_population; returns the previous value"""
        return int(lib().sac_engine_set_grade(self.h, int(grade)))

    def grade_stats(self):
        out = (C.c_longlong * 4)()
        lib().sac_engine_grade_stats(self.h, out)
        return list(out)

    @staticmethod
    def dedup_totals():
        out = (C.c_longlong * 3)()
        lib().sac_dedup_totals(out)
        return list(out)

    def fp64_peak_gflops(self):
        return float(lib().sac_fp64_peak_gflops(self.h))

    def frame_decode(self, nch, data, cap_samples):
        data = np.ascontiguousarray(data, np.uint8)
        planes = [np.zeros(cap_samples, np.int32) for _ in range(nch)]
        arr = (_i32p * nch)(*[_p(p, _i32p) for p in planes])
        n = C.c_int(0)
        used = lib().sac_frame_decode(self.h, nch, _p(data, _u8p), len(data), arr, cap_samples, C.byref(n))
        if used < 0:
            raise SacError("sac_frame_decode failed (%d): %s" % (used, lib().sac_last_error().decode()))
        return [p[:n.value] for p in planes], int(used)

    def encode_memory(self, cfg, wav_bytes):
        wav = np.frombuffer(wav_bytes, np.uint8)
        cap = len(wav) * 2 + 65536
        out = np.zeros(cap, np.uint8)
        olen = C.c_longlong(0)
        st = FileStats()
        _chk(lib().sac_encode_memory(self.h, C.byref(cfg), _p(wav, _u8p), len(wav), _p(out, _u8p), cap, C.byref(olen), C.byref(st)),
             "sac_encode_memory")
        return out[:olen.value].tobytes(), st

    def decode_memory(self, sac_bytes, cap):
        sac = np.frombuffer(sac_bytes, np.uint8)
        out = np.zeros(cap, np.uint8)
        olen = C.c_longlong(0)
        st = FileStats()
        _chk(lib().sac_decode_memory(self.h, _p(sac, _u8p), len(sac), _p(out, _u8p), cap, C.byref(olen), C.byref(st)), "sac_decode_memory")
        return out[:olen.value].tobytes(), st
