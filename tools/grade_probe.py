"""canonical vs search-grade kernels: device time per kernel class for a default-profile generation and for a heterogeneous
(first-DDS-generation-like) one, and the cost agreement. usage: grade_probe.py [window] [P]"""
import sys, time
sys.path[:0] = [".", "tests", "tools"]
import numpy as np, sac_b200 as sb, oracle_lib as ol
from synth_wav import synth_pcm
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
P = int(sys.argv[2]) if len(sys.argv) > 2 else 128
eng = sb.Engine(0); vmin, vmax, vdef = sb.base_profile()
eng.set_dedup(0)
pcm = synth_pcm(3, 2, 3).astype(np.int32); planes, means, mm = ol.analyse([pcm[:, 0], pcm[:, 1]]); win = eng.window(planes, mm)
idx = np.array(sb.SEARCH_DIMS)
lo, hi = vmin[idx].astype(np.float64), vmax[idx].astype(np.float64)
Xd = np.tile(vdef[idx].astype(np.float64), (P, 1))
rng = np.random.default_rng(5)
Xh = Xd.copy()
for p in range(1, P):
    x = Xh[p] + 0.25 * (hi - lo) * rng.standard_normal(56)
    x = np.where(x < lo, np.minimum(lo + (lo - x), hi), x); x = np.where(x > hi, np.maximum(hi - (x - hi), lo), x)
    Xh[p] = x
for name, X in (("default", Xd), ("heterogeneous", Xh)):
    res = {}
    for grade in (0, 1):
        eng.set_grade(grade)
        eng.eval_population(win, 1000, n, vdef, X, sb.COST_BITPLANE, 4)
        t = time.time(); c = eng.eval_population(win, 1000, n, vdef, X, sb.COST_BITPLANE, 4); dt = time.time() - t
        ms, ln = eng.last_timing()
        res[grade] = c
        print(f"{name} grade {grade}: P={P} n={n} wall {dt:.3f}s ols {ms[3]:.1f} ms cascade {ms[0]-ms[3]:.1f} ms bitplane {ms[1]:.1f} ms  clk/sample: ols {ms[3]*1e-3*1.965e9/n:.0f} cascade {(ms[0]-ms[3])*1e-3*1.965e9/n:.0f}", flush=True)
    fin = np.isfinite(res[0]) & np.isfinite(res[1])
    rel = np.abs(res[1][fin] - res[0][fin]) / res[0][fin] if fin.any() else np.array([np.nan])
    print(f"{name}: finite {fin.sum()}/{P}, equal costs {np.mean(res[1][fin] == res[0][fin]):.2f}, max rel diff {rel.max():.2e}, argmin {np.argmin(res[0])} {np.argmin(res[1])}, stats {eng.grade_stats()}", flush=True)
