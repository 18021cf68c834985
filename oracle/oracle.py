"""ctypes view of oracle/libaid_oracle.so -- TEST INFRASTRUCTURE (see aid_oracle.c header).

Importable only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libaid_oracle.so")

ROW_DTYPE = np.dtype([("count", "<i4"), ("track", "<u4"), ("offset", "<i4"),
                      ("q_first", "<i4"), ("q_last", "<i4")])


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, "aid_oracle.c"), os.path.join(_HERE, "dedup_oracle.c"), os.path.join(_HERE, "aid_cpu_f32.c"),
            os.path.join(_HERE, "..", "include", "aid_params.h")]
    stale = (not os.path.exists(_SO)) or force
    if not stale and all(os.path.exists(s) for s in srcs):
        stale = os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        f32p, u32p, i64p, u64p, u8p = (C.POINTER(t) for t in (C.c_float, C.c_uint32, C.c_int64, C.c_uint64, C.c_uint8))
        L.aid_oracle_window.argtypes = [f32p]
        L.aid_oracle_num_frames.argtypes = [C.c_int64]; L.aid_oracle_num_frames.restype = C.c_int64
        L.aid_oracle_stft.argtypes = [f32p, C.c_int64, f32p]; L.aid_oracle_stft.restype = C.c_int64
        L.aid_oracle_peaks.argtypes = [f32p, C.c_int64, u32p]; L.aid_oracle_peaks.restype = C.c_int64
        L.aid_oracle_hashes.argtypes = [u32p, C.c_int64, u32p, u32p]; L.aid_oracle_hashes.restype = C.c_int64
        L.aid_oracle_fingerprint_batch.argtypes = [f32p, i64p, C.c_int, u32p, u32p, i64p, i64p, i64p, C.c_int]
        L.aid_oracle_fingerprint_batch.restype = C.c_int
        L.aid_oracle_index_build.argtypes = [u32p, u32p, u32p, C.c_int64, u64p]
        L.aid_oracle_match.argtypes = [u32p, u32p, u64p, u8p, u32p, u32p, C.c_int64, C.c_void_p, C.c_int]
        L.aid_oracle_match.restype = C.c_int
        L.aid_oracle_params.argtypes = [C.POINTER(C.c_int32)]
        L.aid_cpu_f32_stft.argtypes = [f32p, C.c_int64, f32p]; L.aid_cpu_f32_stft.restype = C.c_int64
        L.aid_cpu_f32_peaks_from_pcm.argtypes = [f32p, C.c_int64, u32p]; L.aid_cpu_f32_peaks_from_pcm.restype = C.c_int64
        L.aid_cpu_f32_fingerprint_batch.argtypes = [f32p, i64p, C.c_int, u32p, u32p, i64p, i64p, i64p, C.c_int]
        L.aid_cpu_f32_fingerprint_batch.restype = C.c_int
        f64p = C.POINTER(C.c_double)
        L.aid_oracle_fp_similarity.argtypes = [u32p, C.c_int64, u32p, C.c_int64]
        L.aid_oracle_fp_similarity.restype = C.c_double
        L.aid_oracle_dedup_scan.argtypes = [u32p, i64p, f64p, C.c_int64, u32p, i64p, f64p, f64p, C.c_int, i64p, f64p, C.c_int]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


PARAM_NAMES = ["sample_rate", "nfft", "hop", "nbins", "peak_half_f", "peak_half_t", "peak_min_bin",
               "dt_min", "dt_max", "df_min", "df_max", "fanout", "min_votes", "max_rows",
               "query_max_frames", "seg_tracks"]


def params() -> dict:
    out = np.zeros(16, np.int32)
    lib().aid_oracle_params(_p(out, C.c_int32))
    return dict(zip(PARAM_NAMES, (int(v) for v in out)))


def window() -> np.ndarray:
    w = np.zeros(1024, np.float32)
    lib().aid_oracle_window(_p(w, C.c_float))
    return w


def num_frames(n: int) -> int:
    return int(lib().aid_oracle_num_frames(n))


def peak_cap(frames: int) -> int:
    return ((frames + 255) // 256) * 2048


def stft(pcm: np.ndarray) -> np.ndarray:
    pcm = np.ascontiguousarray(pcm, np.float32)
    T = num_frames(len(pcm))
    S = np.zeros((T, 512), np.float32)
    if T:
        lib().aid_oracle_stft(_p(pcm, C.c_float), len(pcm), _p(S, C.c_float))
    return S


def peaks(S: np.ndarray) -> np.ndarray:
    """S[T,512] -> sorted uint32 keys (t << 9 | f). Raises OverflowError past the capacity."""
    S = np.ascontiguousarray(S, np.float32)
    T = S.shape[0]
    cap = peak_cap(T)
    keys = np.zeros(max(1, cap), np.uint32)
    n = lib().aid_oracle_peaks(_p(S, C.c_float), T, _p(keys, C.c_uint32))
    if n < 0:
        raise OverflowError("peak capacity exceeded")
    return keys[:n].copy()


def hashes(keys: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    keys = np.ascontiguousarray(keys, np.uint32)
    h = np.zeros(max(1, len(keys) * 8), np.uint32)
    t = np.zeros_like(h)
    n = lib().aid_oracle_hashes(_p(keys, C.c_uint32), len(keys), _p(h, C.c_uint32), _p(t, C.c_uint32))
    return h[:n].copy(), t[:n].copy()


def fingerprint(pcm: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    return hashes(peaks(stft(pcm)))


def stft_f32(pcm: np.ndarray) -> np.ndarray:
    """The tuned single-precision CPU baseline's spectrogram (aid_cpu_f32.c)."""
    pcm = np.ascontiguousarray(pcm, np.float32)
    T = num_frames(len(pcm))
    S = np.zeros((T, 512), np.float32)
    if T:
        lib().aid_cpu_f32_stft(_p(pcm, C.c_float), len(pcm), _p(S, C.c_float))
    return S


def peaks_f32(pcm: np.ndarray) -> np.ndarray:
    """Peaks of the streaming single-precision baseline, straight from PCM."""
    pcm = np.ascontiguousarray(pcm, np.float32)
    keys = np.zeros(max(1, peak_cap(num_frames(len(pcm)))), np.uint32)
    n = lib().aid_cpu_f32_peaks_from_pcm(_p(pcm, C.c_float), len(pcm), _p(keys, C.c_uint32))
    if n < 0:
        raise OverflowError("peak capacity exceeded")
    return keys[:n].copy()


def fingerprint_batch(pcm: np.ndarray, sample_off: np.ndarray, threads: int = 0, f32: bool = False):
    """Ragged batch -> (hash, t_anchor, hash_off[n+1] dense, n_hash, n_peaks, threads_used).
    f32=True runs the tuned single-precision streaming baseline (aid_cpu_f32.c) instead of the f64 checker."""
    pcm = np.ascontiguousarray(pcm, np.float32)
    sample_off = np.ascontiguousarray(sample_off, np.int64)
    n = len(sample_off) - 1
    frames = np.array([num_frames(int(sample_off[i + 1] - sample_off[i])) for i in range(n)], np.int64)
    caps = ((frames + 255) // 256) * 2048 * 8
    slot = np.zeros(n + 1, np.int64); slot[1:] = np.cumsum(caps)
    h = np.zeros(max(1, int(slot[-1])), np.uint32); t = np.zeros_like(h)
    nh = np.zeros(n, np.int64); npk = np.zeros(n, np.int64)
    fn = lib().aid_cpu_f32_fingerprint_batch if f32 else lib().aid_oracle_fingerprint_batch
    used = fn(_p(pcm, C.c_float), _p(sample_off, C.c_int64), n,
                                              _p(h, C.c_uint32), _p(t, C.c_uint32), _p(slot, C.c_int64),
                                              _p(nh, C.c_int64), _p(npk, C.c_int64), threads)
    off = np.zeros(n + 1, np.int64); off[1:] = np.cumsum(np.maximum(nh, 0))
    hd = np.concatenate([h[slot[i]:slot[i] + max(nh[i], 0)] for i in range(n)]) if n else h[:0]
    td = np.concatenate([t[slot[i]:slot[i] + max(nh[i], 0)] for i in range(n)]) if n else t[:0]
    return hd, td, off, nh, npk, used


class Index:
    """Global (unsegmented) sorted index: the definition the segmented GPU index must agree with."""

    def __init__(self, hash_: np.ndarray, track: np.ndarray, t: np.ndarray):
        self.hash = np.ascontiguousarray(hash_, np.uint32).copy()
        self.track = np.ascontiguousarray(track, np.uint32).copy()
        self.t = np.ascontiguousarray(t, np.uint32).copy()
        self.bucket = np.zeros((1 << 24) + 1, np.uint64)
        lib().aid_oracle_index_build(_p(self.hash, C.c_uint32), _p(self.track, C.c_uint32), _p(self.t, C.c_uint32),
                                     len(self.hash), _p(self.bucket, C.c_uint64))

    def match(self, q_hash: np.ndarray, q_t: np.ndarray, tombstone: np.ndarray | None = None, max_rows: int = 50):
        q_hash = np.ascontiguousarray(q_hash, np.uint32); q_t = np.ascontiguousarray(q_t, np.uint32)
        rows = np.zeros(max_rows, ROW_DTYPE)
        tomb = None if tombstone is None else _p(np.ascontiguousarray(tombstone, np.uint8), C.c_uint8)
        n = lib().aid_oracle_match(_p(self.track, C.c_uint32), _p(self.t, C.c_uint32), _p(self.bucket, C.c_uint64),
                                   tomb, _p(q_hash, C.c_uint32), _p(q_t, C.c_uint32), len(q_hash),
                                   rows.ctypes.data, max_rows)
        return rows[:n].copy()


# ---------------------------------------------------------------- content-duplicate scan (dedup_oracle.c)
def fp_similarity(a: np.ndarray, b: np.ndarray) -> float:
    a = np.ascontiguousarray(a, dtype=np.uint32); b = np.ascontiguousarray(b, dtype=np.uint32)
    return float(lib().aid_oracle_fp_similarity(_p(a, C.c_uint32), a.size, _p(b, C.c_uint32), b.size))


def dedup_scan(words, off, dur, q_words, q_off, q_lo, q_hi, n_threads: int = 0):
    """Best row (-1 if none) and similarity per query; arguments as aid_dedup_add / aid_dedup_scan."""
    words = np.ascontiguousarray(words, dtype=np.uint32); off = np.ascontiguousarray(off, dtype=np.int64)
    dur = np.ascontiguousarray(dur, dtype=np.float64)
    q_words = np.ascontiguousarray(q_words, dtype=np.uint32); q_off = np.ascontiguousarray(q_off, dtype=np.int64)
    q_lo = np.ascontiguousarray(q_lo, dtype=np.float64); q_hi = np.ascontiguousarray(q_hi, dtype=np.float64)
    nq = q_lo.size
    best_row = np.full(nq, -1, dtype=np.int64); best_sim = np.zeros(nq, dtype=np.float64)
    lib().aid_oracle_dedup_scan(_p(words, C.c_uint32), _p(off, C.c_int64), _p(dur, C.c_double), dur.size,
                                _p(q_words, C.c_uint32), _p(q_off, C.c_int64), _p(q_lo, C.c_double),
                                _p(q_hi, C.c_double), nq, _p(best_row, C.c_int64), _p(best_sim, C.c_double), n_threads)
    return best_row, best_sim
