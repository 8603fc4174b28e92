"""one DDS-like generation (candidates perturbed around the default profile the way OptDDS does early in a search):
kernel times for heterogeneous candidates + parity of a few costs against the oracle"""
import sys, time
sys.path[:0] = [".", "tests", "tools"]
import numpy as np, sac_b200 as sb, oracle_lib as ol
from synth_wav import synth_pcm
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
P = int(sys.argv[2]) if len(sys.argv) > 2 else 128
pfrac = float(sys.argv[3]) if len(sys.argv) > 3 else 0.67
ncheck = int(sys.argv[4]) if len(sys.argv) > 4 else 3
eng = sb.Engine(0); vmin, vmax, vdef = sb.base_profile()
eng.set_dedup(0)
pcm = synth_pcm(2, 2, 3).astype(np.int32); planes, means, mm = ol.analyse([pcm[:, 0], pcm[:, 1]]); win = eng.window(planes, mm)
rng = np.random.default_rng(5)
idx = np.array(sb.SEARCH_DIMS)
lo, hi = vmin[idx].astype(np.float64), vmax[idx].astype(np.float64)
X = np.tile(vdef[idx].astype(np.float64), (P, 1))
for p in range(1, P):
    sel = rng.random(56) < pfrac
    x = X[p] + sel * 0.25 * (hi - lo) * rng.standard_normal(56)
    x = np.where(x < lo, np.minimum(lo + (lo - x), hi), x); x = np.where(x > hi, np.maximum(hi - (x - hi), lo), x)
    X[p] = x
def orders(x):
    prof = vdef.copy(); prof[idx] = x.astype(np.float32); r = lambda i: int(round(float(prof[i])))
    return (r(24) + r(9), r(25) + r(26) + abs(r(27)), r(28) + r(29) + r(30) + r(37), r(31) + r(32) + r(33) + r(38))
o = np.array([orders(x) for x in X])
print("n_ols ch0 max %d ch1 max %d mean %.1f; taps ch0 max %d ch1 max %d mean %.0f" % (o[:, 0].max(), o[:, 1].max(), o[:, 1].mean(), o[:, 2].max(), o[:, 3].max(), o[:, 2:].mean()))
for rep in range(2):
    t = time.time(); c = eng.eval_population(win, 0, n, vdef, X, sb.COST_BITPLANE, 4); dt = time.time() - t
    ms, ln = eng.last_timing()
    print(f"P={P} n={n}: wall {dt:.3f}s predictor {ms[0]:.1f} ms bitplane {ms[1]:.1f} ms", flush=True)
bad = 0
for p in list(range(1, 1 + ncheck)) + [int(np.argmax(o[:, 1]))]:
    prof = vdef.copy(); prof[idx] = X[p].astype(np.float32)
    e, _ = ol.oracle_predict(planes, mm, prof, 4, 0, n)
    want = sum(ol.oracle_cost(ol.COST_BITPLANE, e[ch]) for ch in range(2))
    ok = c[p] == want; bad += not ok
    print("cand", p, "orders", o[p], "cost", c[p], want, "OK" if ok else "MISMATCH")
print("PARITY", "OK" if not bad else "FAIL")
