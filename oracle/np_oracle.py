"""Slow, independent numpy/scipy restatement of the same specification (include/aid_params.h), used only to
cross-check oracle/aid_oracle.c in tests/test_oracle.py. TEST INFRASTRUCTURE (see aid_oracle.c header)."""
from __future__ import annotations

import numpy as np

NFFT, HOP, NBINS = 1024, 128, 512
HALF_F, HALF_T, MIN_BIN, MIN_S = 51, 12, 9, np.float32(0.001)
DT_MIN, DT_MAX, DF_MIN, DF_MAX, FANOUT = 2, 33, 1, 128, 8
MIN_VOTES, MAX_ROWS = 6, 50


def window() -> np.ndarray:
    n = np.arange(NFFT, dtype=np.float64)
    return (0.54 - 0.46 * np.cos(2 * np.pi * n / (NFFT - 1))).astype(np.float32)


def stft(pcm: np.ndarray) -> np.ndarray:
    pcm = np.asarray(pcm, np.float32)
    T = 0 if len(pcm) < NFFT else (len(pcm) - NFFT) // HOP + 1
    if T == 0:
        return np.zeros((0, NBINS), np.float32)
    frames = np.lib.stride_tricks.sliding_window_view(pcm, NFFT)[::HOP][:T].astype(np.float64) * window().astype(np.float64)
    X = np.fft.rfft(frames, axis=1)[:, :NBINS]
    return np.log1p(X.real ** 2 + X.imag ** 2).astype(np.float32)


def peaks(S: np.ndarray) -> np.ndarray:
    from scipy.ndimage import maximum_filter
    if S.shape[0] == 0:
        return np.zeros(0, np.uint32)
    M = maximum_filter(S, size=(2 * HALF_T + 1, 2 * HALF_F + 1), mode="constant", cval=-1.0)
    mask = (S == M) & (S > MIN_S)
    mask[:, :MIN_BIN] = False
    t, f = np.nonzero(mask)
    return ((t.astype(np.uint32) << 9) | f.astype(np.uint32)).astype(np.uint32)


def hashes(keys: np.ndarray):
    h, ta = [], []
    t = (keys >> 9).astype(np.int64); f = (keys & 511).astype(np.int64)
    for i in range(len(keys)):
        taken = 0
        for j in range(i + 1, len(keys)):
            dt, df = t[j] - t[i], abs(f[j] - f[i])
            if dt > DT_MAX:
                break
            if dt < DT_MIN or df < DF_MIN or df > DF_MAX:
                continue
            h.append((f[i] << 15) | (f[j] << 6) | dt); ta.append(t[i]); taken += 1
            if taken == FANOUT:
                break
    return np.array(h, np.uint32), np.array(ta, np.uint32)


def match(ix_hash, ix_track, ix_t, q_hash, q_t):
    """rows (count, track, offset, q_first, q_last) ordered (count desc, track, offset), at most MAX_ROWS."""
    votes: dict[tuple[int, int], list[int]] = {}
    by_hash: dict[int, list[int]] = {}
    for p, h in enumerate(ix_hash):
        by_hash.setdefault(int(h), []).append(p)
    for h, tq in zip(q_hash, q_t):
        for p in by_hash.get(int(h), ()):
            votes.setdefault((int(ix_track[p]), int(ix_t[p]) - int(tq)), []).append(int(tq))
    rows = [(len(v), k[0], k[1], min(v), max(v)) for k, v in votes.items() if len(v) >= MIN_VOTES]
    rows.sort(key=lambda r: (-r[0], r[1], r[2]))
    return rows[:MAX_ROWS]


# ---------------------------------------------------------------- decode feed (csrc/resample.cu)
def resample_taps() -> np.ndarray:
    """firwin(61, 1/3, window=("kaiser", 5.0)) in double: the default design of scipy.signal.resample_poly(x, 1, 3)."""
    from scipy.signal import firwin
    return firwin(61, 1.0 / 3.0, window=("kaiser", 5.0))


def resample3(x: np.ndarray) -> np.ndarray:
    """y[j] = sum_m h[m + 30] x[3 j + m], zero outside the clip, j < ceil(n / 3) -- in double, with the float32 taps
    the engine uses (what csrc/resample.cu must reproduce); scipy.signal.resample_poly(x, 1, 3) is the same up to the
    rounding of the taps (tests/test_oracle.py)."""
    x = np.asarray(x, np.float64)
    if len(x) == 0:
        return np.zeros(0, np.float64)
    h = resample_taps().astype(np.float32).astype(np.float64)
    full = np.convolve(x, h)                      # full[i] = sum_k h[k] x[i - k]; y[j] = full[3 j + 30]
    n_out = (len(x) + 2) // 3
    return full[30:30 + 3 * n_out:3][:n_out]
