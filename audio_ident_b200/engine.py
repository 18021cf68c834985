"""Python view of one B200 fingerprint engine (one per process, one process per GPU).

Thin by design: every method is one call into the C ABI (include/audio_ident_b200.h); numpy arrays
are the host buffers, and anything with ``data_ptr()`` (a torch CUDA tensor) can stand in for device
memory. There is no CPU implementation behind any of these methods.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _lib
from ._lib import MATCH_ROW_DTYPE, EngineUnavailable, FpDeviceResult

SAMPLE_RATE = 16000
FRAME_SECONDS = 128 / 16000
MAX_ROWS = 50


class EngineError(RuntimeError):
    def __init__(self, status: int, text: str):
        super().__init__(f"audio_ident_b200: {text} (status {status})")
        self.status = status


def _ptr(a) -> int:
    if a is None:
        return 0
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return int(a.data_ptr())
    return int(a)


def ragged(clips: Sequence) -> tuple[np.ndarray, np.ndarray]:
    """list of float32 arrays / f32le bytes -> (concatenated float32, sample_off int64[n+1])."""
    arrs = [np.frombuffer(c, dtype="<f4") if isinstance(c, (bytes, bytearray, memoryview))
            else np.ascontiguousarray(c, dtype=np.float32).reshape(-1) for c in clips]
    off = np.zeros(len(arrs) + 1, np.int64)
    if arrs:
        off[1:] = np.cumsum([len(a) for a in arrs])
    pcm = np.concatenate(arrs) if arrs else np.zeros(0, np.float32)
    return np.ascontiguousarray(pcm, np.float32), off


class Engine:
    def __init__(self, device: int = 0):
        self._L = _lib.load()
        n = self._L.aid_device_count()
        if n <= 0:
            raise EngineUnavailable("no CUDA device is visible; the fingerprint engine has no CPU fallback")
        h = C.c_void_p()
        rc = self._L.aid_engine_create(int(device), C.byref(h))
        if rc != 0:
            raise EngineUnavailable(f"aid_engine_create(device={device}) failed: {self._L.aid_strerror(rc).decode()}")
        self._h = h
        self.device = int(device)

    # -- lifetime ---------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None):
            for x in self.__dict__.pop("_exchanges", []):
                x.close()
            self._L.aid_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int) -> None:
        if rc != 0:
            text = self._L.aid_strerror(rc).decode()
            if rc == -1:
                text += ": " + self._L.aid_last_error(self._h).decode()
            raise EngineError(rc, text)

    @property
    def launches(self) -> int:
        return int(self._L.aid_launch_count(self._h))

    def sync(self) -> None:
        self._check(self._L.aid_engine_sync(self._h))

    def set_max_batch_frames(self, frames: int) -> None:
        self._check(self._L.aid_engine_set_max_batch_frames(self._h, int(frames)))

    def set_stage_timing(self, on: bool) -> None:
        self._check(self._L.aid_engine_set_stage_timing(self._h, int(bool(on))))

    def set_kernels(self, stft_variant: int = 5, peak_summary: bool = True) -> None:
        """Kernel selection for tests and A/B runs (include/audio_ident_b200.h aid_engine_set_kernels): 0 = the scalar
        FP32 STFT kernel, 5 = the packed one (default; 7 adds the software-pipelined separation); peak_summary = the peak kernel streams the STFT's group maxima.
        Results are bit-identical for every choice."""
        self._check(self._L.aid_engine_set_kernels(self._h, int(stft_variant), int(bool(peak_summary))))

    def stage_times(self) -> dict:
        """{stage: (milliseconds, timed launches)} accumulated since the last call; waits for the work."""
        ms = (C.c_double * 8)()
        n = (C.c_int64 * 8)()
        self._check(self._L.aid_engine_stage_times(self._h, ms, n))
        names = ("stft", "peaks", "compact", "hash", "match", "rank", "index_build", "merge")
        return {name: (float(ms[i]), int(n[i])) for i, name in enumerate(names)}

    def params(self) -> dict:
        out = np.zeros(16, np.int32)
        self._L.aid_get_params(out.ctypes.data_as(C.POINTER(C.c_int32)))
        names = ["sample_rate", "nfft", "hop", "nbins", "peak_half_f", "peak_half_t", "peak_min_bin", "dt_min",
                 "dt_max", "df_min", "df_max", "fanout", "min_votes", "max_rows", "query_max_frames", "seg_tracks"]
        return dict(zip(names, (int(v) for v in out)))

    @staticmethod
    def num_frames(n_samples: int) -> int:
        return 0 if n_samples < 1024 else (n_samples - 1024) // 128 + 1

    @staticmethod
    def _off(a) -> tuple[np.ndarray, "C._Pointer"]:
        a = np.ascontiguousarray(a, np.int64)
        return a, a.ctypes.data_as(C.POINTER(C.c_int64))

    # -- fingerprinting ------------------------------------------------------------------------
    def hash_capacity(self, sample_off: np.ndarray) -> int:
        frames = np.maximum((np.diff(sample_off) - 1024) // 128 + 1, 0)
        return int((((frames + 255) // 256) * 2048 * 8).sum())

    def fingerprint(self, pcm, sample_off, hash_cap: int | None = None):
        """Ragged host batch -> (hash u32, t_anchor u32, hash_off i64[n+1], status i32[n])."""
        sample_off, offp = self._off(sample_off)
        n = len(sample_off) - 1
        if isinstance(pcm, np.ndarray):
            pcm = np.ascontiguousarray(pcm, np.float32)
        cap = self.hash_capacity(sample_off) if hash_cap is None else int(hash_cap)
        h = np.empty(max(cap, 1), np.uint32)
        t = np.empty(max(cap, 1), np.uint32)
        hoff = np.zeros(n + 1, np.int64)
        st = np.zeros(max(n, 1), np.int32)
        self._check(self._L.aid_fingerprint_host(self._h, _ptr(pcm), offp, n, h.ctypes.data, t.ctypes.data, cap,
                                                 hoff.ctypes.data_as(C.POINTER(C.c_int64)),
                                                 st.ctypes.data_as(C.POINTER(C.c_int32))))
        total = int(hoff[n])
        return h[:total], t[:total], hoff, st[:n]

    def fingerprint_into(self, pcm, sample_off, h: np.ndarray, t: np.ndarray, hoff: np.ndarray, st: np.ndarray) -> int:
        """Same, into caller-owned (e.g. pinned) buffers; returns the number of hashes."""
        sample_off, offp = self._off(sample_off)
        n = len(sample_off) - 1
        self._check(self._L.aid_fingerprint_host(self._h, _ptr(pcm), offp, n, _ptr(h), _ptr(t), len(h),
                                                 C.cast(_ptr(hoff), C.POINTER(C.c_int64)),
                                                 C.cast(_ptr(st), C.POINTER(C.c_int32))))
        return int(np.asarray(hoff)[n]) if isinstance(hoff, np.ndarray) else -1

    def fingerprint_dev(self, d_pcm, sample_off, stream=None) -> FpDeviceResult:
        """PCM already on the device; results stay there (engine-owned, valid until the next call)."""
        sample_off, offp = self._off(sample_off)
        res = FpDeviceResult()
        self._check(self._L.aid_fingerprint_dev(self._h, _ptr(d_pcm), offp, len(sample_off) - 1, C.byref(res),
                                                _ptr(stream)))
        return res

    def fingerprint_windows_dev(self, d_pcm, win_begin, win_end, stream=None) -> FpDeviceResult:
        """fingerprint_dev for windows d_pcm[win_begin[i]:win_end[i]] that may overlap (no copies)."""
        b, bp = self._off(win_begin)
        e_, ep = self._off(win_end)
        res = FpDeviceResult()
        self._check(self._L.aid_fingerprint_windows_dev(self._h, _ptr(d_pcm), bp, ep, len(b), C.byref(res), _ptr(stream)))
        return res

    def stft(self, pcm, sample_off) -> np.ndarray:
        sample_off, offp = self._off(sample_off)
        pcm = np.ascontiguousarray(pcm, np.float32)
        frames = int(sum(self.num_frames(int(d)) for d in np.diff(sample_off)))
        spec = np.zeros((frames, 512), np.float32)
        self._check(self._L.aid_stft_host(self._h, _ptr(pcm), offp, len(sample_off) - 1, spec.ctypes.data))
        return spec

    def peaks(self, spec: np.ndarray, frame_off):
        frame_off, offp = self._off(frame_off)
        n = len(frame_off) - 1
        spec = np.ascontiguousarray(spec, np.float32)
        cap = int((((np.diff(frame_off) + 255) // 256) * 2048).sum())
        pk = np.zeros(max(cap, 1), np.uint32)
        poff = np.zeros(n + 1, np.int64)
        st = np.zeros(max(n, 1), np.int32)
        self._check(self._L.aid_peaks_host(self._h, spec.ctypes.data, offp, n, pk.ctypes.data, cap,
                                           poff.ctypes.data_as(C.POINTER(C.c_int64)),
                                           st.ctypes.data_as(C.POINTER(C.c_int32))))
        return pk[:int(poff[n])], poff, st[:n]

    def hashes(self, peaks: np.ndarray, peak_off):
        peak_off, offp = self._off(peak_off)
        n = len(peak_off) - 1
        peaks = np.ascontiguousarray(peaks, np.uint32)
        cap = max(1, len(peaks) * 8)
        h = np.zeros(cap, np.uint32)
        t = np.zeros(cap, np.uint32)
        hoff = np.zeros(n + 1, np.int64)
        self._check(self._L.aid_hashes_host(self._h, peaks.ctypes.data, offp, n, h.ctypes.data, t.ctypes.data, cap,
                                            hoff.ctypes.data_as(C.POINTER(C.c_int64))))
        return h[:int(hoff[n])], t[:int(hoff[n])], hoff

    # -- index --------------------------------------------------------------------------------
    @staticmethod
    def _names(names: Sequence[str]):
        arr = (C.c_char_p * max(len(names), 1))()
        for i, s in enumerate(names):
            arr[i] = str(s).encode()
        return arr

    def index_add(self, pcm, sample_off, names: Sequence[str], device: bool = False) -> np.ndarray:
        sample_off, offp = self._off(sample_off)
        n = len(sample_off) - 1
        if len(names) != n:
            raise ValueError("one name per track")
        if not device and isinstance(pcm, np.ndarray):
            pcm = np.ascontiguousarray(pcm, np.float32)
        ok = np.zeros(max(n, 1), np.uint8)
        fn = self._L.aid_index_add_dev if device else self._L.aid_index_add_host
        self._check(fn(self._h, _ptr(pcm), offp, n, self._names(names), ok.ctypes.data_as(C.POINTER(C.c_uint8))))
        return ok[:n].astype(bool)

    def index_add_fp(self, pcm, sample_off, names: Sequence[str]):
        """index_add on host PCM that also returns the stored fingerprints (for the caller's journal):
        (ok bool[n], hash u32, t_anchor u32, hash_off i64[n+1])."""
        sample_off, offp = self._off(sample_off)
        n = len(sample_off) - 1
        if len(names) != n:
            raise ValueError("one name per track")
        if isinstance(pcm, np.ndarray):
            pcm = np.ascontiguousarray(pcm, np.float32)
        cap = self.hash_capacity(sample_off)
        h = np.empty(max(cap, 1), np.uint32)
        t = np.empty(max(cap, 1), np.uint32)
        hoff = np.zeros(n + 1, np.int64)
        ok = np.zeros(max(n, 1), np.uint8)
        self._check(self._L.aid_index_add_host_fp(self._h, _ptr(pcm), offp, n, self._names(names),
                                                  ok.ctypes.data_as(C.POINTER(C.c_uint8)), h.ctypes.data, t.ctypes.data, cap,
                                                  hoff.ctypes.data_as(C.POINTER(C.c_int64))))
        total = int(hoff[n])
        return ok[:n].astype(bool), h[:total], t[:total], hoff

    def index_add_hashes(self, h, t, hash_off, n_frames, names: Sequence[str]) -> np.ndarray:
        hash_off, offp = self._off(hash_off)
        n_frames, nfp = self._off(n_frames)
        n = len(hash_off) - 1
        h = np.ascontiguousarray(h, np.uint32)
        t = np.ascontiguousarray(t, np.uint32)
        ok = np.zeros(max(n, 1), np.uint8)
        self._check(self._L.aid_index_add_hashes(self._h, h.ctypes.data, t.ctypes.data, offp, nfp, n,
                                                 self._names(names), ok.ctypes.data_as(C.POINTER(C.c_uint8))))
        return ok[:n].astype(bool)

    def index_delete(self, name: str) -> bool:
        rc = self._L.aid_index_delete(self._h, str(name).encode())
        if rc == -5:
            return False
        self._check(rc)
        return True

    def index_commit(self) -> None:
        self._check(self._L.aid_index_commit(self._h))

    def index_set_grouping(self, on: bool) -> None:
        """Full segments share a hash directory eight at a time (default) or keep one table each; rows are the same."""
        self._check(self._L.aid_index_set_grouping(self._h, int(bool(on))))

    def index_clear(self) -> None:
        self._check(self._L.aid_index_clear(self._h))

    def index_stats(self) -> dict:
        out = np.zeros(8, np.int64)
        self._check(self._L.aid_index_stats(self._h, out.ctypes.data_as(C.POINTER(C.c_int64))))
        return {"tracks": int(out[0]), "postings": int(out[1]), "segments": int(out[2]),
                "tracks_total": int(out[3]), "device_bytes": int(out[4]), "segments_grouped": int(out[5])}

    def track_name(self, track: int) -> str:
        buf = C.create_string_buffer(256)
        rc = self._L.aid_index_track_name(self._h, int(track), buf, 256)
        if rc < 0:
            self._check(rc)
        return buf.value.decode()

    def index_save(self, path: str) -> None:
        self._check(self._L.aid_index_save(self._h, str(path).encode()))

    def index_load(self, path: str) -> None:
        self._check(self._L.aid_index_load(self._h, str(path).encode()))

    # -- identification ---------------------------------------------------------------------------
    def query(self, pcm, sample_off, max_rows: int = MAX_ROWS, device: bool = False):
        """Ragged batch of vote windows -> (rows[n, max_rows] MATCH_ROW_DTYPE, n_rows i32[n])."""
        sample_off, offp = self._off(sample_off)
        n = len(sample_off) - 1
        if not device and isinstance(pcm, np.ndarray):
            pcm = np.ascontiguousarray(pcm, np.float32)
        rows = np.zeros((max(n, 1), max_rows), MATCH_ROW_DTYPE)
        nr = np.zeros(max(n, 1), np.int32)
        fn = self._L.aid_query_dev if device else self._L.aid_query_host
        self._check(fn(self._h, _ptr(pcm), offp, n, rows.ctypes.data, max_rows, nr.ctypes.data_as(C.POINTER(C.c_int32))))
        return rows[:n], nr[:n]

    def query_windows(self, pcm, win_begin, win_end, max_rows: int = MAX_ROWS):
        """query() for windows pcm[win_begin[i]:win_end[i]] of a host array that may overlap: the span is uploaded once."""
        b, bp = self._off(win_begin)
        e_, ep = self._off(win_end)
        n = len(b)
        if isinstance(pcm, np.ndarray):
            pcm = np.ascontiguousarray(pcm, np.float32)
        rows = np.zeros((max(n, 1), max_rows), MATCH_ROW_DTYPE)
        nr = np.zeros(max(n, 1), np.int32)
        self._check(self._L.aid_query_windows_host(self._h, _ptr(pcm), bp, ep, n, rows.ctypes.data, max_rows,
                                                   nr.ctypes.data_as(C.POINTER(C.c_int32))))
        return rows[:n], nr[:n]

    def query_hashes(self, h, t, hash_off, max_rows: int = MAX_ROWS):
        hash_off, offp = self._off(hash_off)
        n = len(hash_off) - 1
        h = np.ascontiguousarray(h, np.uint32)
        t = np.ascontiguousarray(t, np.uint32)
        rows = np.zeros((max(n, 1), max_rows), MATCH_ROW_DTYPE)
        nr = np.zeros(max(n, 1), np.int32)
        self._check(self._L.aid_query_hashes(self._h, h.ctypes.data, t.ctypes.data, offp, n, rows.ctypes.data, max_rows,
                                             nr.ctypes.data_as(C.POINTER(C.c_int32))))
        return rows[:n], nr[:n]

    def match_dev(self, d_hash, d_t, d_hash_off, d_hash_len, d_status, n_queries: int, d_rows, d_n_rows,
                  max_rows: int = MAX_ROWS, stream=None) -> None:
        """Device in, device out, asynchronous: see aid_match_dev in include/audio_ident_b200.h."""
        self._check(self._L.aid_match_dev(self._h, _ptr(d_hash), _ptr(d_t), _ptr(d_hash_off), _ptr(d_hash_len),
                                          _ptr(d_status), int(n_queries), _ptr(d_rows), int(max_rows), _ptr(d_n_rows),
                                          _ptr(stream)))

    def exchange(self, rank: int, world: int, max_queries: int, max_hashes_per_rank: int = 0) -> "Exchange":
        """This rank's receive window for sharded identification (aid_exchange_create)."""
        x = Exchange(self, rank, world, max_queries, max_hashes_per_rank)
        self.__dict__.setdefault("_exchanges", []).append(x)
        return x

    def match_exchange_dev(self, xchg: "Exchange", d_hash, d_t, d_hash_off, d_hash_len, d_status, n_queries: int,
                           d_track_map, n_map: int, d_rows, d_n_rows, max_rows: int = MAX_ROWS, stream=None) -> None:
        """aid_match_dev with the ranking kernel storing its rows into every rank's window and a device-side merge:
        d_rows / d_n_rows receive the merged rows of all ranks (see include/audio_ident_b200.h)."""
        self._check(self._L.aid_match_exchange_dev(self._h, xchg._h, _ptr(d_hash), _ptr(d_t), _ptr(d_hash_off),
                                                   _ptr(d_hash_len), _ptr(d_status), int(n_queries), _ptr(d_track_map),
                                                   int(n_map), _ptr(d_rows), int(max_rows), _ptr(d_n_rows), _ptr(stream)))

    def identify_exchange_dev(self, xchg: "Exchange", d_pcm, sample_off, d_track_map, n_map: int, d_rows, d_n_rows,
                              max_rows: int = MAX_ROWS, stream=None) -> None:
        """The whole sharded identification step on the device (aid_identify_exchange_dev): this rank fingerprints
        its slice of the window batch, fingerprints and rows travel through the ranks' windows over NVLink."""
        if isinstance(sample_off, tuple):                   # (win_begin, win_end): windows may overlap
            b, bp = self._off(sample_off[0])
            e_, ep = self._off(sample_off[1])
            self._check(self._L.aid_identify_exchange_windows_dev(self._h, xchg._h, _ptr(d_pcm), bp, ep, len(b),
                                                                  _ptr(d_track_map), int(n_map), _ptr(d_rows), int(max_rows),
                                                                  _ptr(d_n_rows), _ptr(stream)))
            return
        sample_off, offp = self._off(sample_off)
        self._check(self._L.aid_identify_exchange_dev(self._h, xchg._h, _ptr(d_pcm), offp, len(sample_off) - 1,
                                                      _ptr(d_track_map), int(n_map), _ptr(d_rows), int(max_rows),
                                                      _ptr(d_n_rows), _ptr(stream)))

    def identify_exchange_host(self, xchg: "Exchange", pcm, sample_off, d_track_map, n_map: int, rows: np.ndarray,
                               n_rows: np.ndarray, rows_first: int = 0, rows_count: int | None = None,
                               max_rows: int = MAX_ROWS) -> None:
        """aid_identify_exchange_host: host window PCM in (only this rank's slice crosses PCIe), merged rows of windows
        [rows_first, rows_first + rows_count) back in the caller's host arrays (MATCH_ROW_DTYPE [count, max_rows])."""
        if isinstance(sample_off, tuple):                   # (win_begin, win_end): windows may overlap
            b, bp = self._off(sample_off[0])
            e_, ep = self._off(sample_off[1])
            n = len(b)
            cnt = n - rows_first if rows_count is None else int(rows_count)
            self._check(self._L.aid_identify_exchange_windows_host(self._h, xchg._h, _ptr(pcm), bp, ep, n, _ptr(d_track_map),
                                                                   int(n_map), int(rows_first), cnt, _ptr(rows), int(max_rows),
                                                                   _ptr(n_rows)))
            return
        sample_off, offp = self._off(sample_off)
        n = len(sample_off) - 1
        cnt = n - rows_first if rows_count is None else int(rows_count)
        self._check(self._L.aid_identify_exchange_host(self._h, xchg._h, _ptr(pcm), offp, n, _ptr(d_track_map), int(n_map),
                                                       int(rows_first), cnt, _ptr(rows), int(max_rows), _ptr(n_rows)))

    def match_stats(self) -> tuple[int, int]:
        """(query hashes looked up, postings touched) by k_match since the last call, while stage timing was on."""
        out = np.zeros(2, np.int64)
        self._check(self._L.aid_match_stats(self._h, out.ctypes.data_as(C.POINTER(C.c_int64))))
        return int(out[0]), int(out[1])

    # -- decode feed ---------------------------------------------------------------------------
    def resample_48k_to_16k(self, pcm48) -> np.ndarray:
        """float32 48 kHz mono -> float32 16 kHz mono (61-tap zero-phase polyphase decimator on the GPU)."""
        x = np.frombuffer(pcm48, dtype="<f4") if isinstance(pcm48, (bytes, bytearray, memoryview)) else \
            np.ascontiguousarray(pcm48, np.float32)
        out = np.empty(int(self._L.aid_resample_out_len(len(x))), np.float32)
        self._check(self._L.aid_resample_48k_to_16k_host(self._h, _ptr(x), len(x), _ptr(out)))
        return out

    def copy_device(self, d_dst, d_src, nbytes: int, stream=None) -> None:
        self._check(self._L.aid_copy_device(self._h, _ptr(d_dst), _ptr(d_src), int(nbytes), _ptr(stream)))

    # -- raw device memory (bindings without a CUDA runtime of their own) --------------------------------
    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._check(self._L.aid_device_alloc(self._h, int(nbytes), C.byref(p)))
        return int(p.value or 0)

    def device_free(self, ptr: int) -> None:
        self._check(self._L.aid_device_free(self._h, int(ptr)))

    def to_device(self, d_ptr: int, arr: np.ndarray) -> None:
        arr = np.ascontiguousarray(arr)
        self._check(self._L.aid_copy_to_device(self._h, int(d_ptr), arr.ctypes.data, arr.nbytes))

    def to_host(self, d_ptr: int, shape, dtype) -> np.ndarray:
        out = np.empty(shape, dtype)
        self._check(self._L.aid_copy_to_host(self._h, out.ctypes.data, int(d_ptr), out.nbytes))
        return out

    def synth_tracks(self, d_pcm, first_track: int, n_tracks: int, samples_per_track: int, seed: int = 42,
                     stream=None, stride: int = 1) -> None:
        """Track k of the batch is global track first_track + k * stride of the deterministic device corpus."""
        self._check(self._L.aid_synth_tracks_strided_dev(self._h, _ptr(d_pcm), int(first_track), int(stride), int(n_tracks),
                                                         int(samples_per_track), int(seed), _ptr(stream)))


class Exchange:
    """Receive window of one rank for the fused rank + exchange path (include/audio_ident_b200.h, aid_exchange_*)."""

    HANDLE_BYTES = 64

    def __init__(self, engine: Engine, rank: int, world: int, max_queries: int, max_hashes_per_rank: int = 0):
        self.engine, self.rank, self.world, self.max_queries = engine, int(rank), int(world), int(max_queries)
        h = C.c_void_p()
        engine._check(engine._L.aid_exchange_create(engine._h, self.rank, self.world, self.max_queries,
                                                    int(max_hashes_per_rank), C.byref(h)))
        self._h = h

    def handle(self) -> bytes:
        buf = (C.c_uint8 * self.HANDLE_BYTES)()
        self.engine._check(self.engine._L.aid_exchange_handle(self._h, buf))
        return bytes(buf)

    def connect(self, handles: Sequence[bytes]) -> None:
        """handles: one 64-byte handle per rank, in rank order (processes on the same box)."""
        if len(handles) != self.world or any(len(h) != self.HANDLE_BYTES for h in handles):
            raise ValueError("one 64-byte handle per rank")
        blob = (C.c_uint8 * (self.HANDLE_BYTES * self.world)).from_buffer_copy(b"".join(handles))
        self.engine._check(self.engine._L.aid_exchange_connect(self._h, blob))

    def connect_local(self, peers: Sequence["Exchange"]) -> None:
        """peers: the Exchange objects of all ranks, in rank order, living in this process."""
        arr = (C.c_void_p * self.world)(*[p._h for p in peers])
        self.engine._check(self.engine._L.aid_exchange_connect_local(self._h, arr))

    def set_timeout_ms(self, ms: int) -> None:
        self.engine._check(self.engine._L.aid_exchange_set_timeout_ms(self._h, int(ms)))

    def check(self) -> None:
        """Raises if a merge gave up waiting for a peer (synchronises the device)."""
        self.engine._check(self.engine._L.aid_exchange_status(self._h))

    def close(self) -> None:
        if self._h:
            self.engine._L.aid_exchange_destroy(self._h)
            self._h = None
