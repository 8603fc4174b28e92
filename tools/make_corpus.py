#!/usr/bin/env python
"""BASELINE.json configs[3]: the 16-file stereo corpus shape: synth_wav seeds 100..115, durations
10-30 s. The published corpus is 51 014 742 bytes; with 44-byte headers and 4-byte sample-frames the nearest reachable
total is 51 014 740 (12 753 509 sample-frames), which is what this writes.

usage: python tools/make_corpus.py OUTDIR [SCALE]     (SCALE < 1 shortens every file, for probes)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from synth_wav import synth_pcm, write_wav  # noqa: E402

TOTAL_FRAMES = 12753509


def durations(scale=1.0):
    rng = np.random.default_rng(1000)
    w = rng.uniform(10.0, 30.0, 16)
    n = np.floor(w / w.sum() * TOTAL_FRAMES).astype(np.int64)
    n = np.clip(n, 10 * 44100, 30 * 44100)
    n[-1] += TOTAL_FRAMES - n.sum()
    assert n.min() >= 10 * 44100 and n.max() <= 30 * 44100 and n.sum() == TOTAL_FRAMES
    return [max(1, int(x * scale)) for x in n]


def main():
    out = sys.argv[1]
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    os.makedirs(out, exist_ok=True)
    total = 0
    for k, n in enumerate(durations(scale)):
        pcm = synth_pcm(n / 44100.0 + 0.001, 2, 100 + k)[:n]
        p = os.path.join(out, "c4_%02d.wav" % k)
        write_wav(p, pcm)
        total += os.path.getsize(p)
    print("wrote 16 files,", total, "bytes")


if __name__ == "__main__":
    main()
