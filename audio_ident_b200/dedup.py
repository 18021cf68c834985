"""Content-duplicate scan of the reference's ``app.audio.dedup`` on a fingerprint store resident in HBM.

SURVEY.md section 8(f)-4. The reference answers "is this track already ingested?" with a Python loop over every
candidate row (audio-ident-service/app/audio/dedup.py:170-222): parse both comma-separated Chromaprint strings,
XOR / popcount the overlapping prefix, scale by a length penalty (``_fingerprint_similarity``, dedup.py:127-167),
keep the strictly-greater running best, accept it at ``>= threshold``. Here the fingerprints of all ingested
tracks live in one ragged ``uint32`` array on the GPU and a query is one streaming pass of a CUDA kernel
(``csrc/dedup.cu``) that returns the same best row and the same IEEE-double similarity, bit for bit.

=============================  ==========================================  =====================================
name                           reference                                   here
=============================  ==========================================  =====================================
``_fingerprint_similarity``    dedup.py:127-167                            one-row scan (tests / tools)
``check_content_duplicate``    dedup.py:170-222 (async, takes a session)   same signature; scans the attached store
``ContentStore``               the ``tracks`` table columns                ``add`` / ``add_many`` / ``check`` / ``check_many``
                               chromaprint_fingerprint / _duration
=============================  ==========================================  =====================================

``check_file_duplicate``, ``f32le_to_s16le`` and ``generate_chromaprint`` (the ``fpcalc`` subprocess) are not on
this path and stay the reference's. There is no CPU fallback: without the CUDA library every call raises
``DedupUnavailable``.
"""
from __future__ import annotations

import asyncio
import ctypes as C
import threading
import uuid
from typing import Iterable, Sequence

import numpy as np

from . import _lib


class DedupUnavailable(RuntimeError):
    """The CUDA store cannot be used (library not built, no device, or a CUDA failure)."""


def parse_fingerprint(fp: str | None) -> np.ndarray | None:
    """``fpcalc -raw`` text -> uint32 words, or ``None`` where dedup.py:143-147 returns similarity 0.0.

    Uses the reference's own expression (``int(x) for x in fp.split(",")``) so every string it accepts or rejects is
    accepted or rejected here; the words are the integers modulo 2**32, which is all dedup.py:158 looks at.
    """
    if fp is None:
        return None
    try:
        values = [int(x) for x in fp.split(",")]
    except ValueError:
        return None
    if not values:
        return None
    return np.fromiter((v & 0xFFFFFFFF for v in values), dtype=np.uint32, count=len(values))


class ContentStore:
    """Chromaprint fingerprints and durations of the ingested tracks, resident on one GPU.

    Row order is insertion order; on equal similarity the earlier row wins, as the first row the reference's loop
    meets does (dedup.py:206-209).
    """

    def __init__(self, device: int = 0):
        try:
            self._L = _lib.load()
        except _lib.EngineUnavailable as e:
            raise DedupUnavailable(str(e)) from e
        h = C.c_void_p()
        rc = self._L.aid_dedup_create(int(device), C.byref(h))
        if rc != 0:
            raise DedupUnavailable(f"aid_dedup_create(device={device}) failed: {self._L.aid_strerror(rc).decode()} "
                                   "(no CUDA device? there is no CPU fallback)")
        self._h = h
        self._ids: list[uuid.UUID] = []
        self._lock = threading.Lock()

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.aid_dedup_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        return len(self._ids)

    @property
    def launches(self) -> int:
        return int(self._L.aid_dedup_launch_count(self._h))

    @property
    def last_scan_ms(self) -> float:
        return float(self._L.aid_dedup_last_scan_ms(self._h))

    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            raise DedupUnavailable(f"{what}: {self._L.aid_strerror(rc).decode()}: "
                                   f"{self._L.aid_dedup_last_error(self._h).decode()}")

    # ------------------------------------------------------------------ filling
    def add_words(self, track_ids: Sequence[uuid.UUID], words: np.ndarray, word_off: np.ndarray,
                  durations: np.ndarray) -> None:
        """Appends rows given as a ragged uint32 array (bulk load without text parsing)."""
        n = len(track_ids)
        words = np.ascontiguousarray(words, dtype=np.uint32)
        word_off = np.ascontiguousarray(word_off, dtype=np.int64)
        durations = np.ascontiguousarray(durations, dtype=np.float64)
        if word_off.shape != (n + 1,) or durations.shape != (n,) or (n and int(word_off[-1]) != words.size):
            raise ValueError("add_words: inconsistent shapes")
        if n == 0:
            return
        with self._lock:
            first = C.c_int64()
            self._check(self._L.aid_dedup_add(self._h, words.ctypes.data_as(C.c_void_p),
                                              word_off.ctypes.data_as(C.POINTER(C.c_int64)),
                                              durations.ctypes.data_as(C.POINTER(C.c_double)), n, C.byref(first)),
                        "aid_dedup_add")
            assert first.value == len(self._ids)
            self._ids.extend(track_ids)

    def add_many(self, rows: Iterable[tuple[uuid.UUID, str | None, float | None]]) -> int:
        """Appends ``(track_id, chromaprint_fingerprint, chromaprint_duration)`` rows, the columns the reference
        selects (dedup.py:192). Rows the reference's WHERE clause or loop can never match (NULL fingerprint or
        duration, dedup.py:193-194, :204) are skipped; an unparsable fingerprint is stored empty (similarity 0.0).
        Returns the number of rows stored."""
        ids, chunks, offs, durs = [], [], [0], []
        for track_id, fp, dur in rows:
            if fp is None or dur is None:
                continue
            w = parse_fingerprint(fp)
            if w is None:
                w = np.empty(0, dtype=np.uint32)
            ids.append(track_id)
            chunks.append(w)
            offs.append(offs[-1] + w.size)
            durs.append(float(dur))
        if ids:
            self.add_words(ids, np.concatenate(chunks) if chunks else np.empty(0, np.uint32),
                           np.asarray(offs, dtype=np.int64), np.asarray(durs, dtype=np.float64))
        return len(ids)

    def add(self, track_id: uuid.UUID, fingerprint: str | None, duration: float | None) -> bool:
        return self.add_many([(track_id, fingerprint, duration)]) == 1

    # ------------------------------------------------------------------ scanning
    def scan_words(self, q_words: np.ndarray, q_off: np.ndarray, q_lo: np.ndarray, q_hi: np.ndarray
                   ) -> tuple[np.ndarray, np.ndarray]:
        """Raw scan: per query the best row (-1 if none) and its similarity (float64)."""
        nq = len(q_lo)
        q_words = np.ascontiguousarray(q_words, dtype=np.uint32)
        q_off = np.ascontiguousarray(q_off, dtype=np.int64)
        q_lo = np.ascontiguousarray(q_lo, dtype=np.float64)
        q_hi = np.ascontiguousarray(q_hi, dtype=np.float64)
        best_row = np.full(nq, -1, dtype=np.int64)
        best_sim = np.zeros(nq, dtype=np.float64)
        if nq == 0:
            return best_row, best_sim
        with self._lock:
            self._check(self._L.aid_dedup_scan(self._h, q_words.ctypes.data_as(C.c_void_p),
                                               q_off.ctypes.data_as(C.POINTER(C.c_int64)),
                                               q_lo.ctypes.data_as(C.POINTER(C.c_double)),
                                               q_hi.ctypes.data_as(C.POINTER(C.c_double)), nq,
                                               best_row.ctypes.data_as(C.POINTER(C.c_int64)),
                                               best_sim.ctypes.data_as(C.POINTER(C.c_double))), "aid_dedup_scan")
        return best_row, best_sim

    def best_matches(self, queries: Sequence[tuple[str, float]]) -> list[tuple[uuid.UUID | None, float]]:
        """Per ``(fingerprint, duration)`` query: ``(best_match_id, best_similarity)`` as the reference's loop leaves
        them (dedup.py:199-212): ``(None, 0.0)`` if no candidate row has a similarity above zero."""
        out: list[tuple[uuid.UUID | None, float]] = [(None, 0.0)] * len(queries)
        live, chunks, offs, lo, hi = [], [], [0], [], []
        for i, (fp, dur) in enumerate(queries):
            w = parse_fingerprint(fp)
            if w is None:                                  # every similarity is 0.0 (dedup.py:143-151)
                continue
            live.append(i)
            chunks.append(w)
            offs.append(offs[-1] + w.size)
            lo.append(dur * 0.9)                           # dedup.py:189-190, Python floats
            hi.append(dur * 1.1)
        if live and len(self._ids):
            rows, sims = self.scan_words(np.concatenate(chunks), np.asarray(offs, np.int64), np.asarray(lo), np.asarray(hi))
            for i, r, s in zip(live, rows.tolist(), sims.tolist()):
                out[i] = (self._ids[r], s) if r >= 0 else (None, 0.0)
        return out

    def check_many(self, queries: Sequence[tuple[str, float]], threshold: float = 0.85) -> list[uuid.UUID | None]:
        return [tid if (sim >= threshold and tid is not None) else None       # dedup.py:214
                for tid, sim in self.best_matches(queries)]

    def check(self, fingerprint: str, duration: float, threshold: float = 0.85) -> uuid.UUID | None:
        return self.check_many([(fingerprint, duration)], threshold)[0]


# ---------------------------------------------------------------------- module-level mirror of app.audio.dedup
_store: ContentStore | None = None


def attach_store(store: ContentStore | None) -> None:
    """Selects the store ``check_content_duplicate`` scans (one per process; filled from the ``tracks`` table at
    start-up and by ``store.add`` after each successful ingest)."""
    global _store
    _store = store


async def check_content_duplicate(session, fingerprint: str, duration: float, threshold: float = 0.85
                                  ) -> uuid.UUID | None:
    """Signature of dedup.py:170-175. ``session`` is accepted and unused: the candidate rows are already resident.
    The scan runs in a worker thread so the event loop is never blocked."""
    if _store is None:
        raise DedupUnavailable("no ContentStore attached (call audio_ident_b200.dedup.attach_store); there is no CPU fallback")
    return await asyncio.get_running_loop().run_in_executor(None, _store.check, fingerprint, duration, threshold)


def _fingerprint_similarity(fp1: str, fp2: str, device: int = 0) -> float:
    """dedup.py:127-167 for one pair, computed by the scan kernel on a one-row store (parity tests and tools)."""
    a, b = parse_fingerprint(fp1), parse_fingerprint(fp2)
    if a is None or b is None:
        return 0.0
    store = ContentStore(device)
    try:
        store.add_words([uuid.UUID(int=0)], b, np.asarray([0, b.size], np.int64), np.asarray([1.0]))
        _, sims = store.scan_words(a, np.asarray([0, a.size], np.int64), np.asarray([-np.inf]), np.asarray([np.inf]))
        return float(sims[0])
    finally:
        store.close()
