"""which search-grade kernel disagrees with the canonical one, and where (probe). SACB_SG_PARTS selects the parts."""
import sys, os
sys.path[:0] = [".", "tests", "tools"]
import numpy as np, sac_b200 as sb, oracle_lib as ol
from synth_wav import synth_pcm
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
eng = sb.Engine(0); vmin, vmax, vdef = sb.base_profile()
pcm = synth_pcm(1, 2, 8).astype(np.int32); planes, means, mm = ol.analyse([pcm[:, 0], pcm[:, 1]]); win = eng.window(planes, mm)
big = vdef.copy(); big[28] = 6000; big[29] = 1500; big[31] = 3000; big[32] = 900; big[33] = 700; big[38] = 300; big[24] = 30; big[9] = 20; big[25] = 32; big[26] = 30; big[27] = 31
tiny = vdef.copy(); tiny[28] = 256; tiny[29] = 32; tiny[30] = 4; tiny[37] = 2; tiny[31] = 256; tiny[32] = 32; tiny[33] = 4; tiny[38] = 2
profs = [vdef, big, tiny]
for k in (4, 1):
    eng.set_grade(0); c, f0 = eng.predict(win, profs, 100, n, k)
    eng.set_grade(1); s, f1 = eng.predict(win, profs, 100, n, k)
    for p in range(len(profs)):
        for ch in range(2):
            d = s[p, ch].astype(np.int64) - c[p, ch].astype(np.int64)
            nz = np.flatnonzero(d)
            print(f"parts={os.environ.get('SACB_SG_PARTS','3')} k={k} prof {p} ch {ch} flags {f0[p]} {f1[p]} mismatches {len(nz)} first {nz[:5].tolist()} d {d[nz[:5]].tolist()} canon {c[p,ch,nz[:3]].tolist()} sg {s[p,ch,nz[:3]].tolist()}", flush=True)
