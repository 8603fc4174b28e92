"""tests/golden/golden_best_profile.json: the profile the UNMODIFIED reference CLI ended its --best search with on the bench stream's
first frame (5 h 42 min in the build container, profiles/reference_best_bench_frame0_r2.txt: stages of 6348 / 1873 / 335 taps, OLS of 20
and 69 regressors) applied to the first 110 250 sample-frames of that frame by the reference's own FrameCoder (libsacref_nc.so:
Predict() + Encode(), -ffp-contract=off) -- stats, payload lengths and SHA-1s per channel. The restatement must reproduce them
(tests/test_oracle_pin.py::test_reference_best_profile_frame_record): the pin at the parameter ranges a real --best search reaches.

usage (build container, /root/reference present):  python tests/golden/make_golden_best_profile.py /path/to/bench_f0_best.sac"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import oracle_lib as ol
from synth_wav import synth_pcm

N = 110250
sac = open(sys.argv[1], "rb").read()
md = int.from_bytes(sac[18:22], "little")
rec = sac[22 + md + 16:]
prof = np.frombuffer(rec[4:4 + 232], "<f4").copy()
pcm = synth_pcm(60, 2, 3).astype(np.int32)[:N]
s = [np.ascontiguousarray(pcm[:, ch]) for ch in range(2)]
ref = ol.ref_lib(nc=True)
rf = ol.RefFrame(ref, 2, 20 * 44100, optimize=0, cost_kind=ol.COST_BITPLANE)
rf.set_samples(s)
rf.analyse()
ref.ref_frame_set_profile(rf.h, ol._p(prof, ol._f32p))
rf.predict()
rf.encode()
assert np.array_equal(rf.profile(), prof)
out = {"_comment": __doc__.split("\n\n")[0], "seed": 3, "secs": 60, "nch": 2, "n": N,
       "profile_hex": prof.tobytes().hex(), "profile_sha1": hashlib.sha1(prof.tobytes()).hexdigest(),
       "stats": [rf.stats(ch) for ch in range(2)], "payload_len": [], "payload_sha1": []}
for ch in range(2):
    p = rf.encoded(ch)
    out["payload_len"].append(int(len(p)))
    out["payload_sha1"].append(hashlib.sha1(p.tobytes()).hexdigest())
json.dump(out, open(os.path.join(HERE, "golden_best_profile.json"), "w"), indent=1)
print(out["stats"], out["payload_len"])
