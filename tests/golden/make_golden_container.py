"""Generates tests/golden/golden_container.json with the UNMODIFIED reference CLI (oracle/_ref/sac): for WAV images of the
shapes src/file/wav.cpp handles (8/16/24-bit, fmt chunk sizes 16/18/40 incl. WAVE_FORMAT_EXTENSIBLE, extra chunks before
and after 'data', odd-sized chunks, a truncated data chunk, several frames) it records the bytes of the .sac file that
precede the first frame record (header, metadata, MD5) and the samples of every frame record. The WAVs are rebuilt
deterministically by wav_case(); nothing but the JSON is committed. Run in the build container only:

    make -C oracle ref && python tests/golden/make_golden_container.py
"""
import hashlib
import json
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from synth_wav import synth_pcm  # noqa: E402


def _chunk(cid, payload, pad=True):
    b = cid + struct.pack("<I", len(payload)) + payload
    return b + (b"\0" if pad and len(payload) & 1 else b"")


def wav_case(name):
    """-> WAV image (bytes). 16-bit-range synthetic audio re-quantised to the container's width."""
    spec = CASES_ALL[name]
    sr, nch, width, secs = spec["sr"], spec["nch"], spec["width"], spec["secs"]
    x = synth_pcm(secs, min(nch, 2), spec["seed"], sr).astype(np.int32)
    if nch > 2:
        x = np.concatenate([x, x[:, :1]], axis=1)
    if width == 1:
        data = ((x >> 8) + 128).astype(np.uint8).tobytes()
    elif width == 2:
        data = x.astype("<i2").tobytes()
    elif width == 4:
        data = (x << 16).astype("<i4").tobytes()
    else:
        v = (x << 8) + (np.arange(x.size).reshape(x.shape) % 251)
        b = v.astype("<i4").tobytes()
        data = b"".join(b[i:i + 3] for i in range(0, len(b), 4))
    bits = spec.get("bits", 8 * width)
    fmt = struct.pack("<HHIIHH", 0xFFFE if spec["fmt"] == 40 else 1, nch, sr, sr * nch * width, nch * width, spec.get("fmt_bits", 8 * width))
    if spec["fmt"] == 18:
        fmt += struct.pack("<H", 0)
    elif spec["fmt"] == 20:
        fmt += struct.pack("<HH", 2, 0)
    elif spec["fmt"] == 40:
        fmt += struct.pack("<HHI", 22, bits, 3 if nch == 2 else 4) + struct.pack("<H", 1) + bytes.fromhex("000000001000800000aa00389b71")
    body = b"WAVE" + _chunk(b"fmt ", fmt)
    for cid, payload in spec.get("before", []):
        body += _chunk(cid.encode(), payload.encode())
    declared = len(data) + spec.get("declare_extra", 0)
    body += b"data" + struct.pack("<I", declared) + data
    if declared & 1 and not spec.get("declare_extra"):
        body += b"\0"
    for cid, payload in spec.get("after", []):
        body += _chunk(cid.encode(), payload.encode())
    body += spec.get("stray", b"")
    return b"RIFF" + struct.pack("<I", len(body)) + body


CASES = {
    "pcm16_stereo": dict(sr=44100, nch=2, width=2, secs=0.3, seed=71, fmt=16),
    "pcm8_mono": dict(sr=22050, nch=1, width=1, secs=0.4, seed=72, fmt=16),
    "pcm24_stereo_extensible": dict(sr=48000, nch=2, width=3, secs=0.2, seed=73, fmt=40, bits=24),
    "pcm16_mono_fmt18": dict(sr=44100, nch=1, width=2, secs=0.25, seed=74, fmt=18),
    "chunks_before_and_after": dict(sr=32000, nch=2, width=2, secs=0.25, seed=75, fmt=16, before=[("LIST", "INFOabc"), ("fact", "1234")],
                                    after=[("id3 ", "tag-bytes-odd"), ("bext", "even")]),
    "truncated_data": dict(sr=44100, nch=2, width=2, secs=0.2, seed=76, fmt=16, declare_extra=4000),
    "three_frames_8k": dict(sr=8000, nch=1, width=2, secs=45.0, seed=77, fmt=16),
    "empty_data": dict(sr=44100, nch=2, width=2, secs=0.0, seed=78, fmt=16),
    "one_sample": dict(sr=44100, nch=1, width=2, secs=1.0 / 44100, seed=79, fmt=16),
    "zero_length_chunk": dict(sr=44100, nch=2, width=2, secs=0.1, seed=80, fmt=16, before=[("LIST", "")], after=[("junk", "")]),
    "stray_bytes_after_data": dict(sr=44100, nch=1, width=2, secs=0.1, seed=81, fmt=16, stray=b"\x01\x02\x03"),
    "fmt_chunk_of_20_bytes": dict(sr=44100, nch=1, width=2, secs=0.05, seed=82, fmt=20),
    "three_channels": dict(sr=44100, nch=3, width=2, secs=0.05, seed=83, fmt=16),
    "pcm32": dict(sr=44100, nch=2, width=4, secs=0.05, seed=84, fmt=16),
}
# Inputs this implementation refuses although the reference CLI does not stop on them: it prints "error: unknown csize" and
# writes a file whose audio cannot be restored (wav.cpp:93-121 has no branch for these containers). Recorded with what the
# reference did; tests only assert the refusal.
REFUSED_HERE = {
    "pcm24_in_32bit_container": dict(sr=44100, nch=2, width=4, secs=0.05, seed=85, fmt=16, fmt_bits=24),
    "zero_bits": dict(sr=44100, nch=1, width=2, secs=0.05, seed=86, fmt=16, fmt_bits=0),
}
CASES_ALL = dict(CASES, **REFUSED_HERE)


def main():
    sac = os.path.join(ROOT, "oracle", "_ref", "sac")
    assert os.path.exists(sac), "build oracle/_ref first (make -C oracle ref)"
    out = {"generator": "tests/golden/make_golden_container.py", "cases": []}
    with tempfile.TemporaryDirectory() as tmp:
        for name in CASES:
            wav = wav_case(name)
            open(os.path.join(tmp, "a.wav"), "wb").write(wav)
            if os.path.exists(os.path.join(tmp, "a.sac")):
                os.remove(os.path.join(tmp, "a.sac"))
            r = subprocess.run([sac, "--encode", "--normal", "a.wav", "a.sac"], cwd=tmp, capture_output=True, text=True)
            if not os.path.exists(os.path.join(tmp, "a.sac")):          # the reference refused the input
                assert "not a valid .wav file" in r.stdout or "unsupported input format" in r.stderr, (name, r.stdout, r.stderr)
                out["cases"].append(dict(name=name, wav_sha1=hashlib.sha1(wav).hexdigest(), rejected=True))
                print(name, len(wav), "-> rejected by the reference")
                continue
            assert r.returncode == 0, (name, r.stdout, r.stderr)
            img = open(os.path.join(tmp, "a.sac"), "rb").read()
            mds = struct.unpack("<I", img[18:22])[0]
            prefix = img[:22 + mds + 16]
            lst = subprocess.run([sac, "--listfull", "a.sac"], cwd=tmp, capture_output=True, text=True).stdout
            frames = [int(l.split()[2]) for l in lst.splitlines() if l.startswith("Frame ")]
            out["cases"].append(dict(name=name, wav_sha1=hashlib.sha1(wav).hexdigest(), prefix_len=len(prefix), prefix_hex=prefix.hex(),
                                     frames=frames))
            print(name, len(wav), "->", len(img), "prefix", len(prefix), "frames", frames)
        out["refused_here"] = []
        for name in REFUSED_HERE:
            wav = wav_case(name)
            open(os.path.join(tmp, "a.wav"), "wb").write(wav)
            if os.path.exists(os.path.join(tmp, "a.sac")):
                os.remove(os.path.join(tmp, "a.sac"))
            try:
                r = subprocess.run([sac, "--encode", "--normal", "a.wav", "a.sac"], cwd=tmp, capture_output=True, text=True, timeout=60)
                what = "rejected" if not os.path.exists(os.path.join(tmp, "a.sac")) else ("accepted, rc %d, %d bytes" % (r.returncode, os.path.getsize(os.path.join(tmp, "a.sac"))))
                msg = [l for l in (r.stdout + r.stderr).splitlines() if "error" in l.lower() or "unsupported" in l.lower()][:2]
            except subprocess.TimeoutExpired:
                what, msg = "did not terminate within 60 s", []
            out["refused_here"].append(dict(name=name, wav_sha1=hashlib.sha1(wav).hexdigest(), reference=what, reference_messages=msg))
            print(name, len(wav), "-> reference:", what, msg)
    json.dump(out, open(os.path.join(HERE, "golden_container.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
