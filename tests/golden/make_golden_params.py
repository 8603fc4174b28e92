"""Generates tests/golden/golden_params.json from the REFERENCE ITSELF: FrameCoder::SetParam (src/libsac/libsac.cpp:37-92)
for the default profile and seeded random profiles inside the search box (negative nS1 = swapped channel order included),
flattened in the order of sac_profile_params (include/sac_b200.h). Run in the build container only:

    make -C oracle ref && python tests/golden/make_golden_params.py
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402


def profiles(vmin, vmax, vdef):
    out = [vdef.copy()]
    rng = np.random.default_rng(2024)
    for i in range(15):
        u = rng.random(58).astype(np.float32)
        p = (vmin + u * (vmax - vmin)).astype(np.float32)
        if i % 3 == 0:
            p[27] = -abs(p[27]) - 1.0
        if i == 5:
            p = np.where(rng.random(58) < 0.5, vmin, vmax).astype(np.float32)      # corners of the box
        out.append(p)
    return out


def main():
    ref = ol.ref_lib(nc=True)
    assert ref is not None, "build oracle/_ref first (make -C oracle ref)"
    vmin, vmax, vdef = ol.base_profile(ref)
    rf = ol.RefFrame(ref, 2, 20 * 44100)
    out = {"generator": "tests/golden/make_golden_params.py", "cases": []}
    for p in profiles(vmin, vmax, vdef):
        buf = (C.c_double * 64)()
        n = ref.ref_set_param(rf.h, ol._p(np.ascontiguousarray(p, np.float32), ol._f32p), buf)
        out["cases"].append([float(buf[i]).hex() for i in range(n)])
    json.dump(out, open(os.path.join(HERE, "golden_params.json"), "w"))
    print(len(out["cases"]), "profiles,", n, "values each")


if __name__ == "__main__":
    main()
