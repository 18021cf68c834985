"""Shared parity helpers -- TEST INFRASTRUCTURE (same rule as oracle.py: tests/, smoke() and bench.py's checker legs only).

End to end the GPU and the double-precision oracle see spectrograms that differ within AID_SPEC_TOL, and peak picking
compares floats for equality, so their peak sets may differ -- but only at NEAR-TIES: a point one side accepts and the
other rejects must sit within tolerance of its neighbourhood maximum (or of the magnitude gate) on the side that
rejected it. explain_peak_diffs checks exactly that for every one-sided peak and raises AssertionError otherwise.
"""
from __future__ import annotations

import numpy as np

TOL = 1e-4


def explain_peak_diffs(S_ref: np.ndarray, S_gpu: np.ndarray, pk_ref: np.ndarray, pk_gpu: np.ndarray) -> int:
    """Number of peaks present on one side only; AssertionError if any of them is not a near-tie."""
    only = np.setxor1d(pk_ref, pk_gpu)
    if len(only) == 0:
        return 0
    T = S_ref.shape[0]
    for k in only:
        t, f = int(k >> 9), int(k & 511)
        t0, t1, f0, f1 = max(0, t - 12), min(T, t + 13), max(0, f - 51), min(512, f + 52)
        for S in (S_ref, S_gpu):
            m = float(S[t0:t1, f0:f1].max())
            gap = m - float(S[t, f])
            thr_gap = abs(float(S[t, f]) - 0.001)
            assert gap <= 2 * TOL * max(abs(m), 1.0) or thr_gap <= 2 * TOL, (t, f, gap)
    return int(len(only))


def explain_track(engine, oracle_mod, pcm: np.ndarray) -> int:
    """One track through both sides, stage by stage: spectrogram within tolerance, every one-sided peak a near-tie,
    GPU hashes equal to the oracle's hasher on the GPU's own peaks. Returns the number of one-sided peaks."""
    S_ref = oracle_mod.stft(pcm)
    S_gpu = engine.stft(pcm, [0, len(pcm)])
    assert S_gpu.shape == S_ref.shape
    assert (np.abs(S_gpu - S_ref) <= TOL * np.maximum(np.abs(S_ref), 1.0)).all(), "spectrogram outside tolerance"
    pk_ref = oracle_mod.peaks(S_ref)
    pk_gpu, _, st = engine.peaks(S_gpu, [0, S_gpu.shape[0]])
    assert st[0] == 0
    n = explain_peak_diffs(S_ref, S_gpu, pk_ref, pk_gpu)
    h, t, _, st = engine.fingerprint(pcm, [0, len(pcm)])
    gh, gt = oracle_mod.hashes(pk_gpu)
    assert st[0] == 0 and np.array_equal(h, gh) and np.array_equal(t, gt), "fused path differs from its own stages"
    return n
