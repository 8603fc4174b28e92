"""what the bench's search asks of the kernels: OLS order and NLMS tap counts of every candidate of ONE frame's search (a SACB_TRACE_DDS
dump of bench.py), per generation -- which kernel class each chain runs on (ols_warp <= 32 < ols_kernel; cascade_sg small / large).
Host only (sac_profile_params). usage: python tools/search_mix.py gpurun_out/bench_trace_first.jsonl"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sac_b200 as sb

rows = [json.loads(l) for l in open(sys.argv[1])]
rows = [r for r in rows if r["frame_n"] == 882000]
first = []
for r in rows:
    if first and r["step"] <= first[-1]["step"]:
        break
    first.append(r)
_, _, vdef = sb.base_profile()
dims = list(sb.SEARCH_DIMS)
print("%-10s %6s | %-29s | %-33s | %s" % ("steps", "cands", "OLS order (per chain)", "NLMS taps per chain (sum of stages)", "cost"))
print("%-10s %6s | %8s %8s %10s | %8s %8s %8s %6s | %s" % ("", "", "median", "max", "> 32", "median", "p90", "max", ">3500", "best so far"))
best = float("inf")
for g0 in range(0, len(first), 128):
    g = first[g0:g0 + 128]
    orders, taps = [], []
    for r in g:
        p = vdef.copy(); p[dims] = np.asarray(r["x"], np.float32)
        q = sb.profile_params(p)
        nA, nB, nM0, nS0, nS1 = [int(v) for v in q[:5]]
        orders += [nA + nM0, nB + nS0 + nS1]                       # engine.cu ols_order: regressors of chain 0 / chain 1
        vn = q[8:16].reshape(2, 4)
        taps += [int(vn[0].sum()), int(vn[1].sum())]
        best = min(best, r["cost"])
    o = np.array(orders); t = np.array(taps)
    print("%4d..%-4d %6d | %8d %8d %9.0f%% | %8d %8d %8d %5.0f%% | %d" % (g[0]["step"], g[-1]["step"], len(g), np.median(o), o.max(), 100 * np.mean(o > 32),
                                                                   np.median(t), np.percentile(t, 90), t.max(), 100 * np.mean(t > 3500), best))
