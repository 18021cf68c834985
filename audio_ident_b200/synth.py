"""Seeded synthetic audio for parity runs (host side, numpy).

The reference ships no audio fixtures, so inputs are defined here,
following SURVEY.md section 8(d): 16 kHz mono float32 in [-1, 1]; track k uses seed
``BASE_SEED + k`` with ``BASE_SEED = 42`` (the reference's corpus seed,
reference audio-ident-service/scripts/build_eval_corpus.py:51). Content is a sum of
Gaussian-windowed tone bursts over a -50 dBFS white-noise floor, peak-normalised to 0.9.
Queries are excerpts at a random *sample* offset with white Gaussian noise at a given SNR
(the reference's noise model: scripts/build_eval_corpus.py:154-169, default 20 dB :603-606).

bench.py generates its large inputs on the device instead (csrc/synth.cu); this module is
what the tests and the golden fixtures use.
"""
from __future__ import annotations

import numpy as np

SAMPLE_RATE = 16000
BASE_SEED = 42
QUERY_SEED_OFFSET = 10**6


def make_track(k: int, seconds: float, *, base_seed: int = BASE_SEED) -> np.ndarray:
    """Track ``k``: float32 array of ``round(seconds * 16000)`` samples."""
    n = int(round(seconds * SAMPLE_RATE))
    if n <= 0:
        return np.zeros(0, np.float32)
    rng = np.random.Generator(np.random.PCG64(base_seed + k))
    x = np.zeros(n, dtype=np.float64)
    n_bursts = max(1, int(round(40 * seconds)))
    centre = rng.uniform(0.0, n, n_bursts)
    freq = np.exp(rng.uniform(np.log(150.0), np.log(7500.0), n_bursts))
    dur = rng.uniform(0.030, 0.300, n_bursts) * SAMPLE_RATE          # samples, ~ +-2 sigma
    amp = rng.uniform(0.05, 0.5, n_bursts)
    phase = rng.uniform(0.0, 2 * np.pi, n_bursts)
    for c, f, d, a, p in zip(centre, freq, dur, amp, phase):
        sigma = d / 4.0
        lo = max(0, int(c - 3 * sigma))
        hi = min(n, int(c + 3 * sigma) + 1)
        if hi <= lo:
            continue
        t = np.arange(lo, hi, dtype=np.float64)
        x[lo:hi] += a * np.exp(-0.5 * ((t - c) / sigma) ** 2) * np.sin(2 * np.pi * f * t / SAMPLE_RATE + p)
    x += rng.standard_normal(n) * 10 ** (-50 / 20)
    peak = np.max(np.abs(x))
    if peak > 0:
        x *= 0.9 / peak
    return x.astype(np.float32)


def make_query(track: np.ndarray, q: int, seconds: float = 5.0, snr_db: float = 20.0,
               *, base_seed: int = BASE_SEED) -> tuple[np.ndarray, int]:
    """Query ``q``: a noisy excerpt of ``track``. Returns (pcm float32, start sample)."""
    n = int(round(seconds * SAMPLE_RATE))
    rng = np.random.Generator(np.random.PCG64(base_seed + QUERY_SEED_OFFSET + q))
    start = int(rng.integers(0, max(1, len(track) - n + 1)))
    clip = track[start:start + n].astype(np.float64)
    p_sig = float(np.mean(clip ** 2))
    p_noise = p_sig / (10 ** (snr_db / 10)) if p_sig > 0 else 0.0
    noisy = clip + rng.standard_normal(len(clip)) * np.sqrt(p_noise)
    return np.clip(noisy, -1.0, 1.0).astype(np.float32), start
