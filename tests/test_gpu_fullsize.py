"""-m gpu: BASELINE.json-sized inputs through size-independent properties (round trip, idempotence, cost identities)."""
import io
import wave

import numpy as np
import pytest

import oracle_lib as ol
import sac_b200 as sb
from synth_wav import synth_pcm

pytestmark = pytest.mark.gpu


def _wav(pcm):
    b = io.BytesIO()
    with wave.open(b, "wb") as w:
        w.setnchannels(pcm.shape[1]); w.setsampwidth(2); w.setframerate(44100); w.writeframes(pcm.astype("<i2").tobytes())
    return b.getvalue()


def test_config1_mono10_normal_roundtrip_and_size(engine):
    """configs[0]: mono 10 s --normal. Reference: 620 239 bytes; tolerance 0.1 %"""
    wav = _wav(synth_pcm(10, 1, 1))
    sac, st = engine.encode_memory(sb.make_cfg("normal"), wav)
    assert abs(len(sac) - 620239) <= 620, len(sac)
    back, st2 = engine.decode_memory(sac, len(wav) + 64)
    assert st2.md5_ok == 1 and back == wav


def test_stereo60_full_frames_roundtrip(engine):
    """3 full 20-s stereo frames (882 000 samples each): encode -> decode is bit exact; reference --normal size 7 030 207"""
    wav = _wav(synth_pcm(60, 2, 3))
    sac, st = engine.encode_memory(sb.make_cfg("normal"), wav)
    assert st.nframes == 3
    assert abs(len(sac) - 7030207) <= 7100, len(sac)
    back, st2 = engine.decode_memory(sac, len(wav) + 64)
    assert st2.md5_ok == 1 and back == wav


def test_best_window_population_properties(engine):
    """--best objective at full window size (441 000 x 2): identical candidates give identical costs, the cost of the
    start vector equals the byte count of the payloads the final coder emits for the same residuals"""
    pcm = synth_pcm(20, 2, 3).astype(np.int32)
    planes, means, mm = ol.analyse([pcm[:, 0], pcm[:, 1]])
    win = engine.window(planes, mm)
    _, _, vdef = sb.base_profile()
    x0 = vdef[sb.SEARCH_DIMS].astype(np.float64)
    X = np.stack([x0, x0, x0 * 1.0])
    c = engine.eval_population(win, 220500, 441000, vdef, X, sb.COST_BITPLANE, 4)
    assert c[0] == c[1] == c[2]
    res, _ = engine.predict(win, [vdef], 220500, 441000, 4)
    nb = sum(len(engine.bitplane_encode(res[0, ch])[0]) for ch in range(2))
    assert nb == c[0]
    win.close()
