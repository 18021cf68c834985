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
There is no CPU fallback: if the CUDA engine cannot be opened every entry point raises ``OlafError``.
"""
from __future__ import annotations

import asyncio
import logging
import os
import struct
import threading
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


def _replay_journal(eng, path: Path) -> int:
    if not path.exists():
        return 0
    data = path.read_bytes()
    pos, n = 0, 0
    while pos + 24 <= len(data):
        magic, kind, name_len, n_frames, n_hash = struct.unpack_from("<4sIIqI", data, pos)
        if magic != _JOURNAL_MAGIC:
            break
        end = pos + 24 + name_len + (8 * n_hash if kind == _JOURNAL_ADD else 0)
        if end > len(data):
            break                                     # torn tail: the last call did not complete
        name = data[pos + 24:pos + 24 + name_len].decode()
        if kind == _JOURNAL_ADD:
            arr = np.frombuffer(data, dtype="<u4", count=2 * n_hash, offset=pos + 24 + name_len)
            eng.index_add_hashes(arr[:n_hash], arr[n_hash:], [0, n_hash], [n_frames], [name])
        elif kind == _JOURNAL_DEL:
            eng.index_delete(name)
        pos = end
        n += 1
    return n


def _append_journal(kind: int, name: str, n_frames: int = 0, h: np.ndarray | None = None, t: np.ndarray | None = None) -> None:
    d = _state.dir
    if d is None:
        return
    nb = name.encode()
    n_hash = 0 if h is None else len(h)
    with open(_journal_path(d), "ab") as f:
        f.write(struct.pack("<4sIIqI", _JOURNAL_MAGIC, kind, len(nb), int(n_frames), n_hash))
        f.write(nb)
        if n_hash:
            f.write(np.ascontiguousarray(h, "<u4").tobytes())
            f.write(np.ascontiguousarray(t, "<u4").tobytes())
        _state.journal_bytes = f.tell()
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
        _state.engine.index_save(str(_state.dir))
        jp = _journal_path(_state.dir)
        if jp.exists():
            jp.unlink()
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
        h, t, hoff, st = eng.fingerprint(pcm, off)
        names = [str(items[i][1]) for i in live]
        frames = [_frames(int(off[k + 1] - off[k])) for k in range(len(live))]
        good = [k for k in range(len(live)) if not (st[k] & 3)]
        if good:
            sel_h = np.concatenate([h[hoff[k]:hoff[k + 1]] for k in good]) if good else h[:0]
            sel_t = np.concatenate([t[hoff[k]:hoff[k + 1]] for k in good]) if good else t[:0]
            sel_off = np.concatenate([[0], np.cumsum([hoff[k + 1] - hoff[k] for k in good])])
            ok = eng.index_add_hashes(sel_h, sel_t, sel_off, [frames[k] for k in good], [names[k] for k in good])
            for j, k in enumerate(good):
                if ok[j]:
                    _append_journal(_JOURNAL_ADD, names[k], frames[k], h[hoff[k]:hoff[k + 1]], t[hoff[k]:hoff[k + 1]])
                    out[live[k]] = True
                else:
                    logger.error("engine refused track %s (longer than the index limit?)", names[k])
        for k in range(len(live)):
            if st[k] & 3:
                logger.error("fingerprinting failed for track %s (status %d)", names[k], int(st[k]))
    return out


def query_many_sync(clips: Sequence[bytes]) -> list[list[OlafMatch]]:
    from .engine import ragged

    out: list[list[OlafMatch]] = [[] for _ in clips]
    live = [i for i, c in enumerate(clips) if c]
    if not live:
        return out
    eng = get_engine()
    with _state.lock:
        windows, owner, starts = [], [], []
        max_samples = (32768 - 1) * 128 + 1024            # AID_QUERY_MAX_FRAMES per vote window
        for i in live:
            arr = np.frombuffer(clips[i], dtype="<f4")
            for s in range(0, max(len(arr), 1), max_samples):
                windows.append(arr[s:s + max_samples]); owner.append(i); starts.append(s // 128)
        pcm, off = ragged(windows)
        rows, n = eng.query(pcm, off)
        for w, i in enumerate(owner):
            for r in rows[w][:n[w]]:
                name = eng.track_name(int(r["track"]))
                q0, q1 = int(r["q_first"]) + starts[w], int(r["q_last"]) + starts[w]
                off_f = int(r["offset"]) - starts[w]
                out[i].append(OlafMatch(int(r["count"]), q0 * FRAME_SECONDS, q1 * FRAME_SECONDS, name, int(r["track"]),
                                        (q0 + off_f) * FRAME_SECONDS, (q1 + off_f) * FRAME_SECONDS))
    for lst in out:
        lst.sort(key=lambda m: m.match_count, reverse=True)   # stable: engine order breaks ties
    return out


def delete_track_sync(track_id: uuid.UUID) -> bool:
    eng = get_engine()
    with _state.lock:
        found = eng.index_delete(str(track_id))
        if found:
            _append_journal(_JOURNAL_DEL, str(track_id))
        return found


async def _in_thread(fn, *args):
    return await asyncio.get_running_loop().run_in_executor(None, fn, *args)


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
        return (await _in_thread(query_many_sync, [pcm_16k_f32le]))[0]
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
