"""throughput probe: population sizes on a fixed window; per-kernel durations come from an ncu launch list of this script"""
import sys, time
sys.path[:0] = [".", "tests", "tools"]
import numpy as np, sac_b200 as sb, oracle_lib as ol
from synth_wav import synth_pcm
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
eng = sb.Engine(0); vmin, vmax, vdef = sb.base_profile()
eng.set_dedup(0)
pcm = synth_pcm(2, 2, 3).astype(np.int32); planes, means, mm = ol.analyse([pcm[:, 0], pcm[:, 1]]); win = eng.window(planes, mm)
for P in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "64,128,256,512,768".split(","))]:
    X = np.tile(np.asarray(vdef, np.float64)[sb.SEARCH_DIMS], (P, 1))
    t = time.time(); c = eng.eval_population(win, 0, n, vdef, X, sb.COST_BITPLANE, 4); dt = time.time() - t
    ms, ln = eng.last_timing()
    print(f"P={P} chains={2*P} n={n}: wall {dt:.3f}s predictor {ms[0]:.1f} ms bitplane {ms[1]:.1f} ms -> {2*P*n/(ms[0]+ms[1])/1e3:.1f} M chain-samples/s", flush=True)
