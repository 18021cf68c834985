"""Drop-in for the reference's ``app.audio.fingerprint`` (audio-ident-service/app/audio/fingerprint.py).

Same public names, signatures and return conventions, so ``app.ingest.pipeline`` (pipeline.py:167-173) and
``app.search.exact`` (exact.py:150-191) keep working unchanged:

=====================  ================================================  ==========================
name                   reference                                         here
=====================  ================================================  ==========================
``OlafError``          fingerprint.py:26-27                              same meaning: engine missing / crashed
``OlafMatch``          fingerprint.py:30-50 (7 fields, seconds)          same dataclass
``olaf_index_track``   fingerprint.py:87-155  (``olaf_c store``)         Engine.fingerprint + index_add_hashes
``olaf_query``         fingerprint.py:158-219 (``olaf_c query``)         Engine.query
``olaf_delete_track``  fingerprint.py:222-270 (``olaf_c del``)           Engine.index_delete
``_parse_olaf_*``      fingerprint.py:273-350 (CSV grammar)              same grammar (used by the B2 shim / tests)
=====================  ================================================  ==========================

What changes is only how the engine is reached: a ctypes call into libaudioident_b200.so (dispatched to a
worker thread so the event loop is never blocked -- ctypes releases the GIL) instead of a temp file plus
``fork/exec`` of ``olaf_c``. The index lives in HBM and is persisted under ``settings.olaf_lmdb_path``
(fingerprint.py:79-84) as a snapshot (``aidx_b200.bin``) plus an append-only journal of the tracks added since
(``journal.bin``); an emptied directory is an empty index (``make rebuild-index``, Makefile:84-93).

Additive batch API: ``index_tracks`` and ``query_many``.

Micro-batching: the unmodified exact lane awaits ``olaf_query`` once per sub-window
(exact.py:150-171) and many requests run concurrently on the event loop, so ``olaf_query`` does not call the engine
itself: it hands its window to a process-wide batcher thread that collects whatever arrives within
``AID_BATCH_WAIT_MS`` (default 2 ms; or until ``AID_BATCH_MAX_WINDOWS``, default 4096) and issues ONE engine call
for all of them -- concurrent coroutines share a GPU pass instead of queueing batch-of-1 launches behind a lock.

There is no CPU fallback: if the CUDA engine cannot be opened every entry point raises ``OlafError``.
"""
from __future__ import annotations

import asyncio
import logging
import os
import struct
import threading
import time
import uuid
from dataclasses import dataclass
from pathlib import Path
from typing import Sequence

import numpy as np

logger = logging.getLogger(__name__)

FRAME_SECONDS = 128 / 16000
_JOURNAL_MAGIC = b"AIDJ"
_JOURNAL_ADD, _JOURNAL_DEL = 1, 2
_CHECKPOINT_BYTES = 256 << 20


class OlafError(Exception):
    """Raised when the fingerprint engine is missing or fails (never for "no match")."""


@dataclass
class OlafMatch:
    """One row of a query result; field order and units as in the reference (seconds)."""

    match_count: int
    query_start: float
    query_stop: float
    reference_path: str
    reference_id: int
    reference_start: float
    reference_stop: float


# ------------------------------------------------------------------------------------ configuration
def _index_dir() -> Path:
    """settings.olaf_lmdb_path when running inside the service, else $OLAF_DB / $OLAF_LMDB_PATH / the default."""
    try:
        from app.settings import settings  # type: ignore

        return Path(settings.olaf_lmdb_path)
    except Exception:
        return Path(os.environ.get("OLAF_DB") or os.environ.get("OLAF_LMDB_PATH") or "./data/olaf_db")


# ------------------------------------------------------------------------------------------ engine
class _State:
    def __init__(self) -> None:
        self.lock = threading.RLock()        # writers are serialised here as well as by the callers
        self.engine = None
        self.dir: Path | None = None
        self.journal_bytes = 0


_state = _State()


def _journal_path(d: Path) -> Path:
    return d / "journal.bin"


class _DirLock:
    """Exclusive advisory lock on the index directory for the duration of a write (journal append, checkpoint): the
    reference's LMDB is single-writer (fingerprint.py:7-8) and several service processes may share one directory."""

    def __init__(self, d: Path) -> None:
        self.path, self.fd = d / ".lock", None

    def __enter__(self):
        import fcntl
        self.fd = os.open(str(self.path), os.O_CREAT | os.O_RDWR, 0o644)
        fcntl.flock(self.fd, fcntl.LOCK_EX)
        return self

    def __exit__(self, *exc):
        import fcntl
        try:
            fcntl.flock(self.fd, fcntl.LOCK_UN)
        finally:
            os.close(self.fd)


def _replay_journal(eng, path: Path) -> int:
    """Re-applies the journal in order. Consecutive additions go to the engine as one batch (one ctypes call, one
    dirty-segment mark); a deletion or a repeated name flushes the batch so the order of effects is kept. The engine
    validates every record (hash / t_anchor ranges: a damaged file must not reach the device); replay stops at the
    first record that fails or does not parse -- what follows a damaged record cannot be trusted."""
    if not path.exists():
        return 0
    data = path.read_bytes()
    pos, n = 0, 0
    batch: list[tuple[str, int, np.ndarray, np.ndarray]] = []
    names_in_batch: set[str] = set()

    def flush() -> bool:
        if not batch:
            return True
        hs = np.concatenate([b[2] for b in batch]) if batch else np.zeros(0, np.uint32)
        ts = np.concatenate([b[3] for b in batch]) if batch else np.zeros(0, np.uint32)
        off = np.concatenate([[0], np.cumsum([len(b[2]) for b in batch])])
        try:
            eng.index_add_hashes(hs, ts, off, [b[1] for b in batch], [b[0] for b in batch])
            good = True
        except Exception as exc:                                    # find the first bad record of this batch
            good = False
            for name, n_frames, h, t in batch:
                try:
                    eng.index_add_hashes(h, t, [0, len(h)], [n_frames], [name])
                except Exception:
                    logger.error("journal %s: record for track %s is damaged (%s); replay stops there", path, name, exc)
                    break
        batch.clear(); names_in_batch.clear()
        return good

    while pos + 24 <= len(data):
        magic, kind, name_len, n_frames, n_hash = struct.unpack_from("<4sIIqI", data, pos)
        if magic != _JOURNAL_MAGIC or kind not in (_JOURNAL_ADD, _JOURNAL_DEL) or name_len > 4096:
            logger.error("journal %s: unreadable record at byte %d; replay stops there", path, pos)
            break
        end = pos + 24 + name_len + (8 * n_hash if kind == _JOURNAL_ADD else 0)
        if end > len(data):
            break                                     # torn tail: the last call did not complete
        try:
            name = data[pos + 24:pos + 24 + name_len].decode()
        except UnicodeDecodeError:
            logger.error("journal %s: unreadable track name at byte %d; replay stops there", path, pos)
            break
        if kind == _JOURNAL_ADD:
            arr = np.frombuffer(data, dtype="<u4", count=2 * n_hash, offset=pos + 24 + name_len)
            if name in names_in_batch or len(batch) >= 4096:
                if not flush():
                    return n
            batch.append((name, int(n_frames), arr[:n_hash], arr[n_hash:]))
            names_in_batch.add(name)
        else:
            if not flush():
                return n
            eng.index_delete(name)
        pos = end
        n += 1
    flush()
    return n


def _append_journal(kind: int, name: str, n_frames: int = 0, h: np.ndarray | None = None, t: np.ndarray | None = None,
                    sync: bool = True) -> None:
    d = _state.dir
    if d is None:
        return
    nb = name.encode()
    n_hash = 0 if h is None else len(h)
    with _DirLock(d), open(_journal_path(d), "ab") as f:
        f.write(struct.pack("<4sIIqI", _JOURNAL_MAGIC, kind, len(nb), int(n_frames), n_hash))
        f.write(nb)
        if n_hash:
            f.write(np.ascontiguousarray(h, "<u4").tobytes())
            f.write(np.ascontiguousarray(t, "<u4").tobytes())
        _state.journal_bytes = f.tell()
        if sync:
            f.flush()
            os.fsync(f.fileno())                        # the caller is about to report the track as indexed
    if _state.journal_bytes > _CHECKPOINT_BYTES:
        checkpoint()


def get_engine():
    """The process-wide engine, created on first use: opens the GPU, loads snapshot + journal."""
    with _state.lock:
        d = _index_dir()
        if _state.engine is not None and _state.dir == d:
            return _state.engine
        try:
            from .engine import Engine

            if _state.engine is not None:
                _state.engine.close()
                _state.engine = None
            eng = Engine(int(os.environ.get("AID_DEVICE", os.environ.get("LOCAL_RANK", "0"))))
            d.mkdir(parents=True, exist_ok=True)
            eng.index_load(str(d))
            replayed = _replay_journal(eng, _journal_path(d))
            eng.index_commit()
        except OlafError:
            raise
        except Exception as exc:
            raise OlafError(
                f"fingerprint engine not available ({exc}). Build libaudioident_b200.so and make sure a CUDA "
                "device is visible; there is no CPU fallback."
            ) from exc
        _state.engine, _state.dir = eng, d
        jp = _journal_path(d)
        _state.journal_bytes = jp.stat().st_size if jp.exists() else 0
        logger.info("fingerprint engine ready: %s (journal entries replayed: %d)", eng.index_stats(), replayed)
        return eng


def checkpoint() -> None:
    """Write a snapshot of the whole index and drop the journal."""
    with _state.lock:
        if _state.engine is None or _state.dir is None:
            return
        with _DirLock(_state.dir):
            _state.engine.index_save(str(_state.dir))      # data and rename are fsynced before this returns (aid_index_save)
            jp = _journal_path(_state.dir)
            if jp.exists():
                jp.unlink()
                try:
                    fd = os.open(str(_state.dir), os.O_RDONLY)
                    os.fsync(fd); os.close(fd)
                except OSError:
                    pass
        _state.journal_bytes = 0


def shutdown() -> None:
    with _state.lock:
        if _state.engine is not None:
            _state.engine.close()
        _state.engine, _state.dir = None, None


def _frames(n_samples: int) -> int:
    return 0 if n_samples < 1024 else (n_samples - 1024) // 128 + 1


# ------------------------------------------------------------------------------- synchronous cores
def index_tracks_sync(items: Sequence[tuple[bytes, uuid.UUID]]) -> list[bool]:
    from .engine import ragged

    out = [False] * len(items)
    live = [i for i, (pcm, _) in enumerate(items) if pcm]
    for i, (pcm, tid) in enumerate(items):
        if not pcm:
            logger.warning("Empty PCM data provided for indexing track %s", tid)
    if not live:
        return out
    eng = get_engine()
    with _state.lock:
        pcm, off = ragged([items[i][0] for i in live])
        names = [str(items[i][1]) for i in live]
        frames = [_frames(int(off[k + 1] - off[k])) for k in range(len(live))]
        # one engine call: fingerprint, store from the device buffers, and the stored fingerprints back for the journal
        ok, h, t, hoff = eng.index_add_fp(pcm, off, names)
        last = max((k for k in range(len(live)) if ok[k]), default=-1)
        for k in range(len(live)):
            if ok[k]:
                _append_journal(_JOURNAL_ADD, names[k], frames[k], h[hoff[k]:hoff[k + 1]], t[hoff[k]:hoff[k + 1]],
                                sync=(k == last))
                out[live[k]] = True
            else:
                logger.error("engine refused track %s (fingerprinting failed or longer than the index limit)", names[k])
    return out


_MAX_WINDOW_SAMPLES = (32768 - 1) * 128 + 1024            # AID_QUERY_MAX_FRAMES frames per vote window


def query_windows_sync(clips: Sequence[bytes], windows: Sequence[tuple[int, int, int]]) -> list[list[OlafMatch]]:
    """One engine call for windows (clip index, first sample, one-past-last sample) over a batch of clips. Windows of one
    clip may overlap -- the exact lane's three sub-windows of a 5 s clip (exact.py:48-52) -- and every clip crosses PCIe
    once (aid_query_windows_host). Result i is what ``olaf_query(clip[first:last])`` returns: times relative to the
    window, rows sorted by match_count descending."""
    out: list[list[OlafMatch]] = [[] for _ in windows]
    arrs = [np.frombuffer(c, dtype="<f4") if c else np.zeros(0, "<f4") for c in clips]
    base = np.zeros(len(arrs) + 1, np.int64)
    if arrs:
        base[1:] = np.cumsum([len(a) for a in arrs])
    begin, end, owner, starts = [], [], [], []
    for w, (ci, lo, hi) in enumerate(windows):
        lo = max(0, min(int(lo), len(arrs[ci]))); hi = max(lo, min(int(hi), len(arrs[ci])))
        if hi <= lo:
            continue
        for s0 in range(lo, hi, _MAX_WINDOW_SAMPLES):        # a window longer than the vote-window limit is cut into several
            begin.append(base[ci] + s0); end.append(base[ci] + min(s0 + _MAX_WINDOW_SAMPLES, hi))
            owner.append(w); starts.append((s0 - lo) // 128)
    if not begin:
        return out
    eng = get_engine()
    with _state.lock:
        pcm = np.concatenate(arrs).astype(np.float32, copy=False)
        rows, n = eng.query_windows(pcm, np.asarray(begin, np.int64), np.asarray(end, np.int64))
        names: dict[int, str] = {}
        for k, w in enumerate(owner):
            for r in rows[k][:n[k]]:
                tr = int(r["track"])
                name = names.get(tr)
                if name is None:
                    name = names[tr] = eng.track_name(tr)
                q0, q1 = int(r["q_first"]) + starts[k], int(r["q_last"]) + starts[k]
                off_f = int(r["offset"]) - starts[k]
                out[w].append(OlafMatch(int(r["count"]), q0 * FRAME_SECONDS, q1 * FRAME_SECONDS, name, tr,
                                        (q0 + off_f) * FRAME_SECONDS, (q1 + off_f) * FRAME_SECONDS))
    for lst in out:
        lst.sort(key=lambda m: m.match_count, reverse=True)   # stable: engine order breaks ties
    return out


def query_many_sync(clips: Sequence[bytes]) -> list[list[OlafMatch]]:
    """Independent result lists for a batch of clips (each one whole window), one GPU pass."""
    return query_windows_sync(clips, [(i, 0, len(c) // 4) for i, c in enumerate(clips)])


def delete_track_sync(track_id: uuid.UUID) -> bool:
    eng = get_engine()
    with _state.lock:
        found = eng.index_delete(str(track_id))
        if found:
            _append_journal(_JOURNAL_DEL, str(track_id))
        return found


async def _in_thread(fn, *args):
    return await asyncio.get_running_loop().run_in_executor(None, fn, *args)


class _QueryBatcher:
    """Collects the windows of concurrent olaf_query coroutines into one engine call.

    A single daemon thread owns the query side of the engine: it sleeps until a window arrives, keeps collecting for at
    most `wait_s` (or until `max_windows`), then runs query_many_sync on the lot and hands every coroutine its own
    list through its event loop. While a GPU pass is in flight the next batch fills up by itself, so under load the
    wait never adds latency; an idle service pays at most `wait_s` (2 ms of a 3 s budget, orchestrator.py:31)."""

    def __init__(self) -> None:
        self.wait_s = float(os.environ.get("AID_BATCH_WAIT_MS", "2")) / 1e3
        self.max_windows = int(os.environ.get("AID_BATCH_MAX_WINDOWS", "4096"))
        self._cv = threading.Condition()
        self._pending: list[tuple[bytes, asyncio.AbstractEventLoop, asyncio.Future]] = []
        self._thread: threading.Thread | None = None
        self.engine_calls = 0                     # statistics (tests, /metrics)
        self.windows = 0

    def submit(self, pcm: bytes) -> "asyncio.Future":
        loop = asyncio.get_running_loop()
        fut = loop.create_future()
        with self._cv:
            if self._thread is None or not self._thread.is_alive():
                self._thread = threading.Thread(target=self._run, name="aid-query-batcher", daemon=True)
                self._thread.start()
            self._pending.append((pcm, loop, fut))
            self._cv.notify()
        return fut

    @staticmethod
    def _deliver(loop, fut, result, exc) -> None:
        def done():
            if fut.cancelled():
                return
            if exc is not None:
                fut.set_exception(exc)
            else:
                fut.set_result(result)
        try:
            loop.call_soon_threadsafe(done)
        except RuntimeError:                      # the caller's loop is gone (request cancelled, loop closed)
            pass

    def _run(self) -> None:
        while True:
            with self._cv:
                while not self._pending:
                    self._cv.wait()
                deadline = time.monotonic() + self.wait_s
                while len(self._pending) < self.max_windows:
                    left = deadline - time.monotonic()
                    if left <= 0:
                        break
                    self._cv.wait(left)
                batch, self._pending = self._pending[:self.max_windows], self._pending[self.max_windows:]
            try:
                results = query_many_sync([b[0] for b in batch])
                self.engine_calls += 1
                self.windows += len(batch)
                for (_, loop, fut), res in zip(batch, results):
                    self._deliver(loop, fut, res, None)
            except BaseException as exc:          # noqa: BLE001 -- every waiter must hear about it
                for _, loop, fut in batch:
                    self._deliver(loop, fut, None, exc)


_batcher = _QueryBatcher()


# --------------------------------------------------------------------------------------- public API
async def olaf_index_track(pcm_16k_f32le: bytes, track_id: uuid.UUID) -> bool:
    """Index a track's fingerprint hashes. ``b""`` -> False without touching the engine; an engine-side
    refusal -> False; engine missing -> OlafError (reference fingerprint.py:87-155)."""
    if not pcm_16k_f32le:
        logger.warning("Empty PCM data provided for indexing track %s", track_id)
        return False
    try:
        return (await _in_thread(index_tracks_sync, [(pcm_16k_f32le, track_id)]))[0]
    except OlafError:
        raise
    except Exception as exc:
        logger.exception("Unexpected error indexing track %s", track_id)
        raise OlafError(f"Failed to index track {track_id}: {exc}") from exc


async def olaf_query(pcm_16k_f32le: bytes) -> list[OlafMatch]:
    """Query the index with a PCM clip; rows sorted by match_count descending, ``[]`` for no match
    (reference fingerprint.py:158-219)."""
    if not pcm_16k_f32le:
        return []
    try:
        return await _batcher.submit(pcm_16k_f32le)
    except OlafError:
        raise
    except Exception as exc:
        logger.exception("Unexpected error during Olaf query")
        raise OlafError(f"Failed to query Olaf: {exc}") from exc


async def olaf_delete_track(track_id: uuid.UUID) -> bool:
    """Remove a track from the index (reference fingerprint.py:222-270). False if it is not there."""
    try:
        return await _in_thread(delete_track_sync, track_id)
    except OlafError:
        raise
    except Exception as exc:
        logger.exception("Unexpected error deleting track %s", track_id)
        raise OlafError(f"Failed to delete track {track_id}: {exc}") from exc


async def index_tracks(items: Sequence[tuple[bytes, uuid.UUID]]) -> list[bool]:
    """Batch form of olaf_index_track: one GPU pass for the whole list."""
    return await _in_thread(index_tracks_sync, list(items))


async def query_windows(clips: Sequence[bytes], windows: Sequence[tuple[int, int, int]]) -> list[list[OlafMatch]]:
    """Batch form for overlapping windows (clip index, first sample, one-past-last sample): see query_windows_sync."""
    return await _in_thread(query_windows_sync, list(clips), list(windows))


async def query_many(clips: Sequence[bytes]) -> list[list[OlafMatch]]:
    """Batch form of olaf_query: independent result lists, one GPU pass (the exact lane's three sub-windows,
    exact.py:150-171, can be issued as one call)."""
    return await _in_thread(query_many_sync, list(clips))


# ------------------------------------------------------- olaf_c CSV grammar (fingerprint.py:273-350)
def _parts_to_match(parts: list[str]) -> OlafMatch | None:
    """Seven string fields -> OlafMatch, None if a numeric field does not parse."""
    try:
        count, q0, q1 = int(parts[0]), float(parts[1]), float(parts[2])
        ref_id, r0, r1 = int(parts[4]), float(parts[5]), float(parts[6])
    except (ValueError, IndexError):
        logger.debug("Failed to parse Olaf output fields: %s", parts)
        return None
    return OlafMatch(count, q0, q1, parts[3], ref_id, r0, r1)


def _parse_olaf_line(line: str) -> OlafMatch | None:
    """One output line: comma-separated first, semicolon-separated as the fallback; needs >= 7 fields."""
    for sep in (",", ";"):
        parts = [p.strip() for p in line.split(sep)]
        if len(parts) >= 7:
            return _parts_to_match(parts)
    logger.debug("Skipping unparseable Olaf output line: %s", line)
    return None


def _parse_olaf_output(stdout: str) -> list[OlafMatch]:
    """All parseable lines, strongest match first."""
    matches = [m for m in (_parse_olaf_line(ln.strip()) for ln in stdout.strip().splitlines() if ln.strip()) if m is not None]
    matches.sort(key=lambda m: m.match_count, reverse=True)
    return matches


def format_olaf_line(m: OlafMatch) -> str:
    """The inverse of _parse_olaf_line (what the olaf_c-compatible shim prints)."""
    return (f"{m.match_count}, {m.query_start:.3f}, {m.query_stop:.3f}, {m.reference_path}, {m.reference_id}, "
            f"{m.reference_start:.3f}, {m.reference_stop:.3f}")
