"""short workload for ncu: one population evaluation on a small window"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import sac_b200 as sb
import oracle_lib as ol
from synth_wav import synth_pcm
P = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
kind = int(sys.argv[3]) if len(sys.argv) > 3 else sb.COST_BITPLANE
eng = sb.Engine(0)
eng.set_dedup(0)
eng.set_grade(int(os.environ.get('SACB_GRADE', '0')))
vmin, vmax, vdef = sb.base_profile()
pcm = synth_pcm(2, 2, 3).astype(np.int32)
planes, means, mm = ol.analyse([pcm[:, 0], pcm[:, 1]])
win = eng.window(planes, mm)
X = np.tile(np.asarray(vdef, np.float64)[sb.SEARCH_DIMS], (P, 1))
c = eng.eval_population(win, 0, n, vdef, X, kind, 4)
print(c[:2], eng.last_timing())
