import sys; sys.path[:0]=[".","tests","tools"]
import numpy as np, sac_b200 as sb, oracle_lib as ol
from synth_wav import synth_pcm
eng=sb.Engine(0); eng.set_dedup(0); vmin,vmax,vdef=sb.base_profile()
pcm=synth_pcm(1,2,3).astype(np.int32); planes,means,mm=ol.analyse([pcm[:,0],pcm[:,1]]); win=eng.window(planes,mm)
for rep in range(2):
    res,fl=eng.predict(win,[vdef]*4,0,30000,4); ms,ln=eng.last_timing()
print("predictor %.1f ms ols %.1f ms cascade %.1f ms -> cascade %.2f us/sample" % (ms[0], ms[3], ms[0]-ms[3], (ms[0]-ms[3])*1e3/30000))
