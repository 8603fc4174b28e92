"""first GPU contact: parity of the CUDA path vs the oracle on small windows + rough timings"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import sac_b200 as sb
import oracle_lib as ol
from synth_wav import synth_pcm

eng = sb.Engine(0)
eng.set_dedup(0)
vmin, vmax, vdef = sb.base_profile()
print("version", sb.lib().sac_version().decode())

def check_predict(nch, secs, seed, profiles, k, frm, n, tag):
    pcm = synth_pcm(secs, nch, seed).astype(np.int32)
    planes, means, mm = ol.analyse([pcm[:, c] for c in range(nch)])
    win = eng.window(planes, mm)
    t = time.time(); res, flags = eng.predict(win, profiles, frm, n, k); dt = time.time() - t
    ms, ln = eng.last_timing()
    ok = True
    for p in range(len(profiles)):
        e, rc = ol.oracle_predict(planes, mm, profiles[p], k, frm, n)
        for ch in range(nch):
            mism = int((e[ch] != res[p, ch]).sum())
            first = int(np.argmax(e[ch] != res[p, ch])) if mism else -1
            if mism: ok = False
            print(f"  {tag} prof{p} ch{ch}: mismatches {mism} first {first} flags {flags[p]} rc {rc} L1 gpu {np.abs(res[p,ch]).mean():.3f} orc {np.abs(e[ch]).mean():.3f}")
    print(f"{tag}: {'OK' if ok else 'FAIL'} wall {dt:.3f}s kernel {ms[0]:.1f} ms for {len(profiles)*nch} chains x {n} samples -> {ms[0]*1e3/n:.2f} us/sample")
    return win, planes, mm, res

rng = np.random.default_rng(7)
def rand_profile(limit=True):
    u = rng.random(58).astype(np.float32)
    p = (vmin + u * (vmax - vmin)).astype(np.float32)
    if limit:
        p[28] = min(p[28], 2000); p[31] = min(p[31], 1500)
    return p

check_predict(1, 2, 1, [vdef], 4, 0, 20000, "mono default k4")
check_predict(1, 2, 1, [vdef], 1, 100, 8000, "mono default k1")
profs = [vdef, rand_profile(), rand_profile(), rand_profile()]
profs[2][27] = -7.0
win, planes, mm, res = check_predict(2, 2, 2, profs, 4, 500, 20000, "stereo mixed k4")
check_predict(2, 1, 5, [rand_profile(False)], 1, 0, 4000, "stereo big k1")

# costs
for kind, name in ((sb.COST_L1, "L1"), (sb.COST_RMS, "rms"), (sb.COST_ENTROPY, "ent"), (sb.COST_GOLOMB, "glb"), (sb.COST_BITPLANE, "bpn")):
    bufs = res[:, :, :].reshape(-1, res.shape[2])
    t = time.time(); g = eng.cost(kind, bufs); dt = time.time() - t
    o = np.array([ol.oracle_cost(kind, b) for b in bufs])
    print(f"cost {name}: max rel diff {np.max(np.abs(g-o)/np.maximum(np.abs(o),1e-30)):.3e} exact {np.array_equal(g,o)} wall {dt:.3f}s", g[:3], o[:3])

# population eval
X = np.stack([np.asarray(p, np.float64)[sb.SEARCH_DIMS] for p in profs])
for kind in (sb.COST_ENTROPY, sb.COST_BITPLANE):
    c = eng.eval_population(win, 500, 20000, vdef, X, kind, 4)
    o = []
    for p in range(len(profs)):
        e, _ = ol.oracle_predict(planes, mm, profs[p], 4, 500, 20000)
        o.append(sum(ol.oracle_cost(kind, e[ch]) for ch in range(2)))
    print("eval_population", kind, c, np.array(o), eng.last_timing())

# bitplane payload parity + frame round trip
pcm = synth_pcm(2, 2, 9).astype(np.int32)
cfg = sb.make_cfg("normal")
t = time.time(); data, prof = eng.frames_encode(cfg, [[pcm[:, 0], pcm[:, 1]]], 20 * 44100); dt = time.time() - t
print("frame encode bytes", len(data), "wall %.3f" % dt, eng.last_timing())
lib = ol.oracle(); lib.saco_set_modes(ol.ORDER_B200, ol.MATH_CANON)
import ctypes as C
cfgv = (C.c_int * 8)(0, 0, 0, 0, 200000, 4, 2, 20 * 44100)
out = np.zeros(len(data) * 2 + 4096, np.uint8)
pr = vdef.copy()
s0 = np.ascontiguousarray(pcm[:, 0]); s1 = np.ascontiguousarray(pcm[:, 1])
nb = lib.saco_encode_frame(2, len(s0), s0.ctypes.data_as(C.c_void_p), s1.ctypes.data_as(C.c_void_p), pr.ctypes.data_as(C.c_void_p), cfgv,
                           out.ctypes.data_as(C.c_void_p), len(out))
print("oracle frame bytes", nb, "identical", nb == len(data) and np.array_equal(out[:nb], data))
t = time.time(); dec, used = eng.frame_decode(2, data, 20 * 44100); dt = time.time() - t
print("decode used", used, "roundtrip", np.array_equal(dec[0], pcm[:, 0]) and np.array_equal(dec[1], pcm[:, 1]), "wall %.3f" % dt, eng.last_timing())

# throughput probe: default-profile population on a 1 s window
pcm = synth_pcm(4, 2, 3).astype(np.int32)
planes, means, mm = ol.analyse([pcm[:, 0], pcm[:, 1]])
win2 = eng.window(planes, mm)
for P in (8, 64, 256):
    Xp = np.tile(np.asarray(vdef, np.float64)[sb.SEARCH_DIMS], (P, 1))
    for kind in (sb.COST_L1, sb.COST_BITPLANE):
        t = time.time(); c = eng.eval_population(win2, 0, 44100, vdef, Xp, kind, 4); dt = time.time() - t
        print(f"P={P} kind={kind} n=44100: wall {dt:.3f}s timing {eng.last_timing()} cost0 {c[0]}")
