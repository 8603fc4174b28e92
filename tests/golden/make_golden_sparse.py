"""Generates tests/golden/golden_sparse.json from the REFERENCE ITSELF (oracle/_ref/libsacref_nc.so): the sparse-PCM
side of a frame -- Remap::Analyse/Map/Unmap, MapEncoder, CalcRemapError and the mapped/unmapped choice of
FrameCoder::EncodeMonoFrame (src/libsac/map.cpp, libsac.cpp:214-298) -- on seeded synthetic inputs whose sample values
occupy only part of the 16-bit grid. Run in the build container only:

    make -C oracle ref && python tests/golden/make_golden_sparse.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402
from synth_wav import synth_pcm  # noqa: E402


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()


def sparse_pcm(kind, secs, nch, seed):
    """int32 [n, nch]; the kinds are the sparse shapes real files show"""
    x = synth_pcm(secs, nch, seed).astype(np.int32)
    if kind == "shift4":        # 12-bit audio in a 16-bit container
        x = (x >> 4) << 4
    elif kind == "times3":      # integer gain: non-power-of-two gaps
        x = (x // 5) * 3
    elif kind == "offset":      # coarse grid off zero, so the removed mean matters
        x = ((x >> 3) << 3) + 5
    elif kind == "holes":       # a few unused values only: cost ratio stays below 1.05, frame stays unmapped
        x = np.where(x % 37 == 0, x + 1, x)
    elif kind == "ch1only":     # one dense channel, one sparse
        x[:, 1] = (x[:, 1] >> 5) << 5
    elif kind == "dense":
        pass
    else:
        raise ValueError(kind)
    return np.clip(x, -32768, 32767).astype(np.int32)


MAPS = [("shift4", 0.2, 1, 31), ("times3", 0.2, 1, 32), ("offset", 0.1, 1, 33), ("dense", 0.3, 1, 34), ("holes", 0.3, 1, 35)]
FRAMES = [("sp_mono_shift4", "shift4", 0.5, 1, 41), ("sp_stereo_times3", "times3", 0.4, 2, 42),
          ("sp_stereo_offset", "offset", 0.4, 2, 43), ("sp_stereo_ch1only", "ch1only", 0.4, 2, 44),
          ("sp_mono_holes", "holes", 0.4, 1, 45), ("sp_stereo_dense", "dense", 0.3, 2, 46)]


def main():
    ref = ol.ref_lib(nc=True)
    assert ref is not None, "build oracle/_ref first (make -C oracle ref)"
    _, _, vdef = ol.base_profile(ref)
    out = {"generator": "tests/golden/make_golden_sparse.py", "map": [], "frame": []}
    for kind, secs, nch, seed in MAPS:
        raw = np.ascontiguousarray(sparse_pcm(kind, secs, nch, seed)[:, 0])
        buf = np.zeros(1 << 16, np.uint8)
        nb = ref.ref_map_encode(ol._p(raw, ol._i32p), len(raw), ol._p(buf, ol._u8p), len(buf))
        rng = np.random.default_rng(seed)
        n = 4000
        pred = rng.integers(-33000, 33000, n).astype(np.int32)
        err = (rng.laplace(0, 90, n)).astype(np.int32)
        m = np.zeros(n, np.int32)
        ref.ref_remap(ol._p(raw, ol._i32p), len(raw), ol._p(pred, ol._i32p), ol._p(err, ol._i32p), n, 0, ol._p(m, ol._i32p))
        # Unmap of ranks that exist: feed Map's own output back
        um = np.zeros(n, np.int32)
        ref.ref_remap(ol._p(raw, ol._i32p), len(raw), ol._p(pred, ol._i32p), ol._p(m, ol._i32p), n, 1, ol._p(um, ol._i32p))
        out["map"].append(dict(kind=kind, secs=secs, nch=nch, seed=seed, nbytes=int(nb), sha1=sha(buf[:nb]), n=n,
                               map_sha1=sha(m), unmap_sha1=sha(um)))
    for name, kind, secs, nch, seed in FRAMES:
        pcm = sparse_pcm(kind, secs, nch, seed)
        rf = ol.RefFrame(ref, nch, 20 * 44100, optimize=0, sparse_pcm=1)
        rf.set_samples([pcm[:, ch] for ch in range(nch)])
        rf.predict(); rf.encode()
        stats = [rf.stats(ch) for ch in range(nch)]
        pay = [rf.encoded(ch) for ch in range(nch)]
        rec = dict(name=name, kind=kind, secs=secs, nch=nch, seed=seed, stats=stats, maxbpn_map=[rf.maxbpn_map(ch) for ch in range(nch)],
                   payload_sha1=[sha(p) for p in pay], payload_len=[int(len(p)) for p in pay])
        # the reference decodes its own record back to the input (fresh FrameCoder, as the decoder would be)
        rd = ol.RefFrame(ref, nch, 20 * 44100, optimize=0, sparse_pcm=1)
        blocks = [(pay[ch], stats[ch][0], stats[ch][1], stats[ch][2], rf.maxbpn_map(ch) if stats[ch][5] else stats[ch][3], stats[ch][5])
                  for ch in range(nch)]
        dec = rd.decode_blocks(len(pcm), rf.profile(), blocks)
        assert all(np.array_equal(dec[ch], pcm[:, ch]) for ch in range(nch)), name
        out["frame"].append(rec)
        print(name, "mapped", [s[5] for s in stats], "bytes", rec["payload_len"])
    with open(os.path.join(HERE, "golden_sparse.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
