"""GPU parity of the sparse-PCM path: used-value map, rank-mapped residuals, the map coder and the
mapped/unmapped choice of a frame record against the CPU restatement (itself pinned to the reference on the same inputs,
tests/test_oracle_pin.py::test_sparse_pcm_*), and the decode direction (map decode, Unmap inside the predictor)."""
import ctypes as C
import json
import os
import sys

import numpy as np
import pytest

import oracle_lib as ol
import sac_b200 as sb

pytestmark = pytest.mark.gpu
FS = 20 * 44100
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden_sparse import sparse_pcm  # noqa: E402


def _oracle_record(nch, raw, sparse):
    lib = ol.oracle()
    lib.saco_set_modes(ol.ORDER_B200, ol.MATH_CANON)
    _, _, vdef = sb.base_profile()
    cfg = (C.c_int * 8)(0, 0, 0, 0, 200000, 4, 2, FS)
    prof = vdef.copy()
    s = [np.ascontiguousarray(p, np.int32) for p in raw]
    out = np.zeros(8 * len(s[0]) + 4096, np.uint8)
    mapped = (C.c_int * 2)(0, 0)
    nb = lib.saco_encode_frame2(nch, len(s[0]), ol._p(s[0], ol._i32p), ol._p(s[1], ol._i32p) if nch > 1 else None, ol._p(prof, ol._f32p),
                                cfg, int(sparse), ol._p(out, ol._u8p), len(out), mapped)
    return out[:nb].copy(), list(mapped)[:nch]


def _block_flags(rec, nch):
    pos, flags = 4 + 58 * 4, []
    for _ in range(nch):
        bs = int.from_bytes(rec[pos:pos + 4].tobytes(), "little")
        flags.append(int(rec[pos + 16]) | (int(rec[pos + 17]) << 8))
        pos += 18 + bs
    return flags


@pytest.mark.parametrize("name", ["sp_mono_shift4", "sp_stereo_offset", "sp_stereo_ch1only", "sp_mono_holes"])
def test_sparse_frame_record_identical_to_oracle_and_roundtrip(engine, name):
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_sparse.json")))
    c = [x for x in g["frame"] if x["name"] == name][0]
    nch = c["nch"]
    pcm = sparse_pcm(c["kind"], c["secs"], nch, c["seed"])
    raw = [np.ascontiguousarray(pcm[:, ch]) for ch in range(nch)]
    rec, _ = engine.frames_encode(sb.make_cfg(None, optimize=0), [raw], FS)
    orec, omapped = _oracle_record(nch, raw, True)
    assert omapped == [st[5] for st in c["stats"]]                  # the same channels the reference codes rank-mapped
    assert [f >> 9 for f in _block_flags(rec, nch)] == omapped
    assert len(rec) == len(orec) and np.array_equal(rec, orec)
    dec, used = engine.frame_decode(nch, rec, FS)
    assert used == len(rec) and all(np.array_equal(dec[ch], raw[ch]) for ch in range(nch))
    lib = ol.oracle(); lib.saco_set_modes(ol.ORDER_B200, ol.MATH_CANON)
    d = [np.zeros(len(raw[0]), np.int32) for _ in range(nch)]
    n_out = C.c_int(0)
    lib.saco_decode_frame(nch, ol._p(rec, ol._u8p), len(rec), ol._p(d[0], ol._i32p), ol._p(d[1], ol._i32p) if nch > 1 else None, C.byref(n_out))
    assert all(np.array_equal(d[ch], raw[ch]) for ch in range(nch))


def test_sparse_pcm_off_and_wide_samples_stay_unmapped(engine):
    pcm = sparse_pcm("shift4", 0.1, 1, 51)
    raw = [np.ascontiguousarray(pcm[:, 0])]
    rec, _ = engine.frames_encode(sb.make_cfg(None, optimize=0, sparse_pcm=0), [raw], FS)
    orec, _ = _oracle_record(1, raw, False)
    assert np.array_equal(rec, orec) and _block_flags(rec, 1)[0] >> 9 == 0
    # 24-bit material: the reference's map stops at +-32768 and its mapped records of wider samples cannot be decoded;
    # such channels are coded unmapped here and round-trip
    wide = [(raw[0] * 64).astype(np.int32)]
    rec, _ = engine.frames_encode(sb.make_cfg(None, optimize=0), [wide], FS)
    assert _block_flags(rec, 1)[0] >> 9 == 0
    dec, used = engine.frame_decode(1, rec, FS)
    assert used == len(rec) and np.array_equal(dec[0], wide[0])


def test_whole_file_with_split_and_mapped_blocks_equals_the_restatement(engine):
    """file level: WAV with a sparse stretch -> adaptive sub-frame split, one rank-mapped block; the .sac image equals the
    host container plan + the restatement's frame records (canonical arithmetic) byte for byte, and decodes to the input.
    The same assembly in the reference's arithmetic is byte-identical to the reference CLI's file
    (tests/test_oracle_pin.py::test_whole_files_are_byte_identical_to_the_reference_cli)."""
    from make_golden_files import wav_of
    from helpers import oracle_file_image
    wav = wav_of("mono_sparse_middle")
    sac, st = engine.encode_memory(sb.make_cfg("normal"), wav)
    want, flags = oracle_file_image(wav, (ol.ORDER_B200, ol.MATH_CANON))
    assert flags == [0, 1, 0, 0] and st.nframes == 4
    assert len(sac) == len(want) and sac == want
    # ... and for this file that is the very file the reference CLI writes (tests/golden/golden_files.json)
    import hashlib
    gold = {f["name"]: f for f in json.load(open(os.path.join(ROOT, "tests", "golden", "golden_files.json")))["files"]}
    from helpers import reference_view
    assert hashlib.sha1(reference_view(sac)).hexdigest() == gold["mono_sparse_middle_normal"]["sac_sha1"]   # but for header byte 17 (arithmetic variant)
    back, st2 = engine.decode_memory(sac, len(wav) + 64)
    assert st2.md5_ok == 1 and back == wav


def test_cli_encode_decode_roundtrip_and_console_layout(tmp_path):
    """the `sac` CLI end to end on the GPU: encode (DDS in generations, the CLI's default), decode, cmp; console lines follow
    CmdLine::Process (cmdline.cpp:245-358)"""
    import re
    import subprocess
    from make_golden_container import wav_case
    cli = os.path.join(ROOT, "sac_b200", "sac")
    wav = wav_case("chunks_before_and_after")                      # stereo 32 kHz, LIST/fact before and id3/bext after 'data'
    (tmp_path / "a.wav").write_bytes(wav)
    r = subprocess.run([cli, "--encode", "--optimize=0.05,9,ent", "a.wav", "a.sac"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    out = r.stdout
    assert "Open: 'a.wav': ok (%d Bytes)" % len(wav) in out and "  32000Hz 16 Bit  Stereo\n  8000 Samples [00:00:00.250]\n" in out
    assert "Create: 'a.sac': ok\n  Profile: mt2 20s ab zero-mean sparse-pcm\n  Optimize: DDS 5.0%,n=9,ent,k=4\n" in out
    assert re.search(r"\n  MD5:     [0-9a-f]+\n\n  %d->\d+=\d+\.\d%% \(\d+\.\d{3} bps\)  \d+\.\d{3}x\n\n  Time:    \[\d\d:\d\d:\d\d\]\n$" % len(wav), out), out
    r = subprocess.run([cli, "--decode", "a.sac", "b.wav"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "Create: 'b.wav': ok\n" in r.stdout and "  Audio MD5: ok\n" in r.stdout, r.stdout + r.stderr
    assert (tmp_path / "b.wav").read_bytes() == wav
