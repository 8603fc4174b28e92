"""-m gpu: several GPUs from the C++ host (sac_encode_file_multi, sac_encode_files, `sac --gpus`): no collective, frames / files
dealt to one engine per device; bytes equal the single-GPU --opt-reset result. With one GPU visible the same entry points run
with one engine (the N = 2 cases are skipped)."""
import os
import subprocess

import numpy as np
import pytest

import sac_b200 as sb
from synth_wav import synth_pcm

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _wav(path, seconds, seed, sr=8000):
    import wave
    pcm = synth_pcm(seconds, 2, seed, sr).astype("<i2")
    with wave.open(str(path), "wb") as w:
        w.setnchannels(2); w.setsampwidth(2); w.setframerate(sr); w.writeframes(pcm.tobytes())
    return path


def _cfg():
    # three 1-s frames of an 8 kHz file, a short bitplane-cost search per frame (the reference's sequential search, speculative batches)
    return sb.make_cfg(None, optimize=1, fraction=0.5, maxnfunc=12, num_threads=0, spec=4, sigma=0.25, cost_kind=sb.COST_BITPLANE, max_framelen=1, grade=1)


def test_file_over_the_engines_equals_the_single_gpu_reset_result(engine, tmp_path):
    ndev = sb.lib().sac_device_count()
    wav = _wav(tmp_path / "a.wav", 3, 41)
    single = sb.make_cfg(None, **{k: getattr(_cfg(), k) for k, _ in sb.Cfg._fields_})
    single.reset = 1; single.frame_parallel = 2
    ref, st0 = engine.encode_memory(single, open(wav, "rb").read())
    assert st0.nframes == 3
    engines = [engine] + [sb.Engine(d) for d in range(1, min(ndev, 2))]
    try:
        st = sb.encode_file_multi(engines, _cfg(), wav, tmp_path / "a.sac")
        got = open(tmp_path / "a.sac", "rb").read()
        assert st.nframes == 3 and got == ref, (len(got), len(ref), len(engines))
        back, st2 = engine.decode_memory(got, os.path.getsize(wav) + 64)
        assert st2.md5_ok == 1 and back == open(wav, "rb").read()
    finally:
        for e in engines[1:]:
            e.close()


def test_batch_of_files_over_the_engines(engine, tmp_path):
    ndev = sb.lib().sac_device_count()
    wavs = [_wav(tmp_path / ("f%d.wav" % i), 1 + (i % 2), 50 + i) for i in range(4)]
    outs = [tmp_path / ("f%d.sac" % i) for i in range(4)]
    engines = [engine] + [sb.Engine(d) for d in range(1, min(ndev, 2))]
    try:
        stats, status = sb.encode_files(engines, _cfg(), wavs, outs)
        assert status == [0, 0, 0, 0]
        for w, o, st in zip(wavs, outs, stats):
            one, _ = engine.encode_memory(sb.make_cfg(None, **dict({k: getattr(_cfg(), k) for k, _ in sb.Cfg._fields_}, reset=1, frame_parallel=2)), open(w, "rb").read())
            assert open(o, "rb").read() == one and st.out_bytes == len(one)
        with pytest.raises(sb.SacError, match="same output"):
            sb.encode_files(engines, _cfg(), wavs[:2], [outs[0], outs[0]])
    finally:
        for e in engines[1:]:
            e.close()


def test_cli_gpus_flag(tmp_path):
    ndev = sb.lib().sac_device_count()
    if ndev < 2:
        pytest.skip("one GPU visible: `sac --gpus=2` needs two")
    wav = _wav(tmp_path / "a.wav", 3, 43)
    cli = os.path.join(ROOT, "sac_b200", "sac")
    args = ["--encode", "--optimize=0.5,12,bpn", "--framelen=1", "--opt-reset"]
    subprocess.run([cli] + args + ["--frame-parallel=2", "a.wav", "one.sac"], cwd=tmp_path, check=True, capture_output=True, timeout=300)
    r = subprocess.run([cli] + args + ["--gpus=2", "a.wav", "two.sac"], cwd=tmp_path, check=True, capture_output=True, text=True, timeout=300)
    assert "2 GPUs" in r.stdout
    assert open(tmp_path / "one.sac", "rb").read() == open(tmp_path / "two.sac", "rb").read()
