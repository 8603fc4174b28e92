import sys; sys.path[:0]=[".","tests","tools"]
import numpy as np, sac_b200 as sb, oracle_lib as ol
from synth_wav import synth_pcm
eng=sb.Engine(0); vmin,vmax,vdef=sb.base_profile()
pcm=synth_pcm(1,2,3).astype(np.int32); planes,means,mm=ol.analyse([pcm[:,0],pcm[:,1]]); win=eng.window(planes,mm)
for k in (4,1):
    res,fl=eng.predict(win,[vdef],0,16000,k); print("k",k,eng.last_timing())
p2=vdef.copy(); p2[25]=30; p2[26]=20; p2[27]=14   # n1 = 64
res,fl=eng.predict(win,[p2],0,8000,4); print("n64",eng.last_timing())
