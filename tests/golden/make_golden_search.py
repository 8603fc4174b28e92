"""Generates tests/golden/golden_search.json from the REFERENCE ITSELF (oracle/_ref/libsacref_nc.so): the profile that
FrameCoder::Predict() ends with after a DDS search that really moves (tens of the 56 dimensions change), population and
sequential. tests/test_host_logic.py drives the PRODUCT's search driver (sac_dds_run: start vector and first generation as
one batch) with the restatement's objective and must arrive at the same 58 floats. Run in the build container only:

    make -C oracle ref && python tests/golden/make_golden_search.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "tools"))
import oracle_lib as ol  # noqa: E402
from synth_wav import synth_pcm  # noqa: E402

CASES = [
    dict(name="stereo_pop8_l1", nch=2, secs=0.5, seed=61, cfg=dict(fraction=0.004, maxnfunc=41, num_threads=8, sigma=0.25, cost_kind=0)),
    dict(name="mono_seq_ent", nch=1, secs=0.5, seed=62, cfg=dict(fraction=0.006, maxnfunc=30, num_threads=0, sigma=0.2, cost_kind=2)),
    dict(name="stereo_pop16_bpn", nch=2, secs=0.4, seed=63, cfg=dict(fraction=0.003, maxnfunc=50, num_threads=16, sigma=0.25, cost_kind=4)),
]


def main():
    ref = ol.ref_lib(nc=True)
    assert ref is not None, "build oracle/_ref first (make -C oracle ref)"
    _, _, vdef = ol.base_profile(ref)
    out = {"generator": "tests/golden/make_golden_search.py", "cases": []}
    for c in CASES:
        pcm = synth_pcm(c["secs"], c["nch"], c["seed"]).astype(np.int32)
        rf = ol.RefFrame(ref, c["nch"], 20 * 44100, optimize=1, **c["cfg"])
        rf.set_samples([np.ascontiguousarray(pcm[:, ch]) for ch in range(c["nch"])])
        rf.predict()
        p = rf.profile()
        out["cases"].append(dict(c, changed=int(np.sum(p != vdef)), profile_sha1=hashlib.sha1(p.tobytes()).hexdigest()))
        print(c["name"], "changed", out["cases"][-1]["changed"])
    json.dump(out, open(os.path.join(HERE, "golden_search.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
