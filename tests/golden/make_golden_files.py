"""Generates tests/golden/golden_files.json: WHOLE .sac files written by the unmodified reference CLI built with
-ffp-contract=off (oracle/_ref/sac_nc, see oracle/Makefile) -- length and SHA-1 -- for WAVs with sparse stretches (adaptive
sub-frame split, rank-mapped blocks), without and with DDS searches (sequential and population, warm start from frame to
frame). tests/test_oracle_pin.py rebuilds each file from the host container plan + the CPU restatement's frame records
and compares byte for byte. Run in the build container only:

    make -C oracle ref && python tests/golden/make_golden_files.py
"""
import hashlib
import io
import json
import os
import subprocess
import sys
import tempfile
import wave

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_golden_split import CASES as SPLIT_CASES, case_pcm  # noqa: E402
from synth_wav import synth_pcm  # noqa: E402

# name -> (pcm source, CLI flags, the same settings for the restatement)
FILES = [
    dict(name="mono_sparse_middle_normal", src="mono_sparse_middle", flags=["--normal"], optimize=None),
    dict(name="stereo_sparse_tail_normal", src="stereo_sparse_tail", flags=["--normal"], optimize=None),
    dict(name="mono_all_sparse_normal", src="mono_all_sparse", flags=["--normal"], optimize=None),
    dict(name="stereo_sparse_tail_dds9_ent", src="stereo_sparse_tail", flags=["--optimize=0.03,9,ent"],
         optimize=dict(fraction=0.03, maxnfunc=9, sigma=0.2, cost_kind=2)),
    dict(name="mono_alternating_dds7_bpn_nosparse", src="mono_alternating", flags=["--optimize=0.02,7,bpn", "--sparse-pcm=0"],
         optimize=dict(fraction=0.02, maxnfunc=7, sigma=0.2, cost_kind=4), sparse=0),
    dict(name="stereo1s_dds10_pop4", src=("synth", 1.0, 2, 91, 44100), flags=["--optimize=0.02,10,ent", "--opt-cfg=dds,4,0.25"],
         optimize=dict(fraction=0.02, maxnfunc=10, sigma=0.25, cost_kind=2, num_threads=4)),
]


def wav_of(src):
    if isinstance(src, (tuple, list)):
        _, secs, nch, seed, sr = src
        pcm = synth_pcm(secs, nch, seed, sr)
    else:
        name, nch, sr, secs, seed, sparse = [c for c in SPLIT_CASES if c[0] == src][0]
        pcm = case_pcm(nch, sr, secs, seed, sparse)
    b = io.BytesIO()
    with wave.open(b, "wb") as w:
        w.setnchannels(pcm.shape[1]); w.setsampwidth(2); w.setframerate(sr); w.writeframes(pcm.astype("<i2").tobytes())
    return b.getvalue()


def main():
    sac = os.path.join(ROOT, "oracle", "_ref", "sac_nc")
    assert os.path.exists(sac), "build oracle/_ref first (make -C oracle ref)"
    out = {"generator": "tests/golden/make_golden_files.py", "reference": "slmdev/sac v0.7.25, -O3 -march=x86-64-v3 -ffp-contract=off", "files": []}
    with tempfile.TemporaryDirectory() as tmp:
        for f in FILES:
            wav = wav_of(f["src"])
            open(os.path.join(tmp, "a.wav"), "wb").write(wav)
            subprocess.run([sac, "--encode"] + f["flags"] + ["a.wav", "a.sac"], cwd=tmp, check=True, capture_output=True)
            img = open(os.path.join(tmp, "a.sac"), "rb").read()
            out["files"].append(dict(name=f["name"], wav_sha1=hashlib.sha1(wav).hexdigest(), sac_len=len(img), sac_sha1=hashlib.sha1(img).hexdigest()))
            print(f["name"], len(wav), "->", len(img))
    json.dump(out, open(os.path.join(HERE, "golden_files.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
