"""probe: one edge input of tests/test_gpu_grade.py::test_search_grade_edge_inputs. usage: sg_edge.py NAME K [PARTS via SACB_SG_PARTS]"""
import sys
sys.path[:0] = [".", "tests", "tools"]
import numpy as np, sac_b200 as sb, oracle_lib as ol
from synth_wav import synth_pcm
name, k = sys.argv[1], int(sys.argv[2])
sq = (np.where((np.arange(4000) // 50) % 2 == 0, 32767, -32768)).astype(np.int32)
cases = {"n1": [np.array([123], np.int32)], "n3": [np.array([5, -4, 9], np.int32)] * 2, "silence": [np.zeros(3000, np.int32)] * 2,
         "dc": [np.full(2000, -7, np.int32)], "fullscale": [sq, -sq], "ragged": [synth_pcm(1, 1, 3).astype(np.int32)[:2999, 0]]}
eng = sb.Engine(0); _, _, vdef = sb.base_profile()
planes, means, mm = ol.analyse(cases[name]); win = eng.window(planes, mm); n = len(planes[0])
eng.set_grade(1)
fast, fl = eng.predict(win, [vdef], 0, n, k)
e, rc = ol.oracle_predict(planes, mm, vdef, k, 0, n)
print(name, k, "flags", fl, [int(np.count_nonzero(fast[0, ch].astype(np.int64) - e[ch])) for ch in range(len(planes))], flush=True)
