"""sequential (warm start, final pass overlapped) vs per-frame streams (frame_parallel=2): wall time for a small 3-frame stream"""
import sys, time
sys.path[:0] = [".", "tests", "tools"]
import numpy as np, sac_b200 as sb
from synth_wav import synth_pcm
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
nfunc = int(sys.argv[2]) if len(sys.argv) > 2 else 257
nfr = int(sys.argv[3]) if len(sys.argv) > 3 else 3
fs = int(secs * 44100)
eng = sb.Engine(0)
pcm = synth_pcm(secs * nfr, 2, 3).astype(np.int32)
frames = [[np.ascontiguousarray(pcm[f * fs:(f + 1) * fs, 0]), np.ascontiguousarray(pcm[f * fs:(f + 1) * fs, 1])] for f in range(nfr)]
for mode in (0, 2, 0, 2):
    cfg = sb.make_cfg("best", num_threads=128, maxnfunc=nfunc, frame_parallel=mode, max_framelen=int(secs))
    t = time.time(); rec, prof = eng.frames_encode(cfg, frames, fs); dt = time.time() - t
    pos = 0; ok = True
    for fr in frames:
        dec, used = eng.frame_decode(2, rec[pos:], fs); ok &= all(np.array_equal(dec[ch], fr[ch]) for ch in range(2)); pos += used
    print(f"mode {mode}: {dt:.2f} s for {nfr} frames of {fs} samples, {len(rec)} bytes, roundtrip {ok and pos == len(rec)}, launches {eng.launches}", flush=True)
