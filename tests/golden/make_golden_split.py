"""Golden for the adaptive sub-frame split (Codec::Analyse, src/libsac/libsac.cpp:726-780): WAVs with sparse stretches
(values on a coarse grid) are encoded by the UNMODIFIED reference CLI (oracle/_ref/sac --normal) and the frame lengths
it chose are read back with --listfull.  usage: python tests/golden/make_golden_split.py  -> tests/golden/golden_split.json"""
import json, os, re, subprocess, sys, tempfile, wave
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.path.insert(0, os.path.join(ROOT, "tools"))
from synth_wav import synth_pcm

# the fixtures are attenuated by 8 so that a 3-s block at 8 kHz (24000 samples) uses its value range densely: at full scale every
# block of so few samples already looks "sparse" to the reference's test (cost ratio > 1.35) and nothing is ever split
CASES = [  # name, channels, sample rate, seconds, seed, sparse stretches [(from_s, to_s, grid)]
    ("mono_sparse_middle", 1, 8000, 21.0, 41, [(6.0, 15.0, 32)]),
    ("stereo_sparse_tail", 2, 8000, 26.5, 42, [(12.0, 26.5, 16)]),
    ("mono_all_sparse", 1, 8000, 9.0, 43, [(0.0, 9.0, 64)]),
    ("stereo_none", 2, 8000, 10.0, 44, []),
    ("mono_alternating", 1, 8000, 40.0, 45, [(3.0, 6.0, 32), (9.0, 12.0, 32), (20.0, 29.0, 8), (31.0, 33.0, 32)]),
    ("stereo_short_sparse_head", 2, 8000, 7.0, 46, [(0.0, 2.0, 32)]),
]


def case_pcm(nch, sr, secs, seed, sparse):
    pcm = synth_pcm(secs, nch, seed, sr).astype(np.int32) // 8
    for a, b, grid in sparse:
        i0, i1 = int(a * sr), int(b * sr)
        pcm[i0:i1] = (pcm[i0:i1] // grid) * grid
    return pcm


def main():
    out = {"generator": "tests/golden/make_golden_split.py", "cases": []}
    ref = os.path.join(ROOT, "oracle", "_ref", "sac")
    with tempfile.TemporaryDirectory() as tmp:
        for name, nch, sr, secs, seed, sparse in CASES:
            pcm = case_pcm(nch, sr, secs, seed, sparse)
            wav = os.path.join(tmp, name + ".wav")
            with wave.open(wav, "wb") as w:
                w.setnchannels(nch); w.setsampwidth(2); w.setframerate(sr); w.writeframes(pcm.astype("<i2").tobytes())
            subprocess.run([ref, "--encode", "--normal", wav, wav + ".sac"], check=True, capture_output=True)
            txt = subprocess.run([ref, "--listfull", wav + ".sac"], check=True, capture_output=True, text=True).stdout
            frames = [int(m) for m in re.findall(r"Frame \d+: (\d+) samples", txt)]
            sparse_flags = [int(m) for m in re.findall(r"sparse_pcm: (\d+)", txt)]
            out["cases"].append(dict(name=name, nch=nch, sr=sr, secs=secs, seed=seed, sparse=sparse, frames=frames,
                                     sparse_pcm_flags=sparse_flags))
            print(name, frames, sparse_flags)
    json.dump(out, open(os.path.join(HERE, "golden_split.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
