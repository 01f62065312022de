"""Deterministic synthetic PCM16 for tests and bench (SURVEY.md §8d / BASELINE.md §3).  Data generation only.

Per channel: three sines (60 Hz..12 kHz, random phase) whose amplitudes sum to a per-channel peak drawn from
{2000, 8000, 20000, 32767}, white noise at -30 dBFS relative to that peak, and one 50 ms full-scale burst per second
(forces the clamp / scale-retry paths of the encoder).  Channels 0..3 of every batch are degenerate on purpose:
all-zero, a full-scale square wave at Nyquist/4, the reference test suite's 261.63 Hz sine
(src/VGAudio.Tests/GenerateAudio.cs:14,23-33) and its ascending ramp (:109-117).
"""
from __future__ import annotations

import numpy as np

SEED = 0x5647415544494F  # "VGAUDIO"
PEAKS = (2000.0, 8000.0, 20000.0, 32767.0)


def reference_sine(n: int, frequency: float, sample_rate: int) -> np.ndarray:
    """GenerateAudio.GenerateSineWave: (short)(short.MaxValue * Math.Sin(c * i)), truncation toward zero."""
    c = 2 * np.pi * frequency / sample_rate
    return np.trunc(32767.0 * np.sin(c * np.arange(n, dtype=np.float64))).astype(np.int16)


def reference_ramp(start: int, count: int) -> np.ndarray:
    """GenerateAudio.GenerateAscendingShorts: pcm[i] = (short)(i + 1 + start)."""
    return (np.arange(count, dtype=np.int64) + 1 + start).astype(np.int16)


def channel(index: int, n: int, sample_rate: int = 48000, degenerate: bool = True) -> np.ndarray:
    if degenerate and index == 0:
        return np.zeros(n, dtype=np.int16)
    if degenerate and index == 1:
        return np.where((np.arange(n) // 4) % 2 == 0, 32767, -32768).astype(np.int16)
    if degenerate and index == 2:
        return reference_sine(n, 261.63, sample_rate)
    if degenerate and index == 3:
        return reference_ramp(0, n)
    rng = np.random.default_rng([SEED, index])
    t = np.arange(n, dtype=np.float64) / sample_rate
    peak = PEAKS[int(rng.integers(0, 4))]
    amps = rng.dirichlet(np.ones(3)) * peak
    x = np.zeros(n, dtype=np.float64)
    for a in amps:
        f = float(np.exp(rng.uniform(np.log(60.0), np.log(12000.0))))
        x += a * np.sin(2 * np.pi * f * t + rng.uniform(0, 2 * np.pi))
    x += rng.normal(0.0, peak * 10 ** (-30 / 20), n)
    burst = int(0.05 * sample_rate)
    for sec in range(0, max(n // sample_rate, 1)):
        start = sec * sample_rate + int(rng.integers(0, max(sample_rate - burst, 1)))
        stop = min(start + burst, n)
        if start < n:
            x[start:stop] = rng.choice(np.array([-32768.0, 32767.0]), stop - start)
    return np.clip(np.rint(x), -32768, 32767).astype(np.int16)


def batch(n_channels: int, n: int, sample_rate: int = 48000, first_index: int = 0, degenerate: bool = True) -> np.ndarray:
    """[n_channels, n] int16, channel-major contiguous."""
    out = np.empty((n_channels, n), dtype=np.int16)
    for c in range(n_channels):
        out[c] = channel(first_index + c, n, sample_rate, degenerate)
    return out
