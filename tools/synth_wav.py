"""Synthetic PCM fixtures: sine mixture + AM tone + white noise, int16 WAV.

usage: python tools/synth_wav.py out.wav SECONDS CHANNELS SEED
"""
import sys
import wave

import numpy as np


def synth_pcm(secs: float, ch: int, seed: int, sr: int = 44100) -> np.ndarray:
    """Returns int16 array [n, ch]."""
    rng = np.random.default_rng(seed)
    n = int(round(secs * sr))
    t = np.arange(n) / sr
    chans = []
    for c in range(ch):
        x = (0.35 * np.sin(2 * np.pi * (220.0 * (c + 1)) * t)
             + 0.2 * np.sin(2 * np.pi * (1333.7 + 17 * c) * t + 0.3)
             + 0.1 * np.sin(2 * np.pi * 3.1 * t) * np.sin(2 * np.pi * 5210.3 * t))
        x += 0.02 * rng.standard_normal(n)
        chans.append(x)
    if ch == 2:
        chans[1] = 0.6 * chans[0] + 0.4 * chans[1]
    pcm = np.clip(np.round(np.stack(chans, axis=1) * 32767 * 0.8), -32768, 32767).astype('<i2')
    return pcm


def write_wav(path: str, pcm: np.ndarray, sr: int = 44100) -> None:
    with wave.open(path, 'wb') as w:
        w.setnchannels(pcm.shape[1])
        w.setsampwidth(2)
        w.setframerate(sr)
        w.writeframes(pcm.tobytes())


if __name__ == '__main__':
    path, secs, ch, seed = sys.argv[1], float(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    write_wav(path, synth_pcm(secs, ch, seed))
