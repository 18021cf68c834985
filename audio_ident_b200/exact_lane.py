"""Batched form of the exact lane's engine-facing half.

The reference's ``run_exact_lane`` (audio-ident-service/app/search/exact.py:70-124) slices a short clip into three
sub-windows and awaits one ``olaf_query`` per window, one clip at a time (exact.py:150-171). Here the windows
of a whole batch of clips go to the GPU in ONE ``query_many`` call; everything after the rows come back is
the reference's arithmetic, restated:

* duration and window slicing in bytes            exact.py:361-399, windows :48-52
* <= 5 s -> 3 windows + consensus, > 5 s -> 1 query  exact.py:100-103
* consensus: >= 2 windows -> sum, 1 window -> max(sum // 2, 1); offset = median(reference_start)   :220-293
* full clip: sum per track, offset = median                                                          :296-332
* drop < 8 aligned hashes, confidence = min(n / 20, 1), sort by confidence desc, cut to max_results  :109-121

The PostgreSQL metadata join (exact.py:407-494) stays in the service. tests/test_boundary_contract.py replays
golden vectors recorded from the reference's own functions (tests/golden/make_golden.py).
"""
from __future__ import annotations

import statistics
import uuid
from dataclasses import dataclass
from typing import Callable, Sequence

from .fingerprint import OlafMatch

MIN_ALIGNED_HASHES = 8
STRONG_MATCH_HASHES = 20
SHORT_CLIP_THRESHOLD_SEC = 5.0
SUB_WINDOWS = [(0.0, 3.5), (0.75, 4.25), (1.5, 5.0)]
SAMPLE_RATE = 16000
BYTES_PER_SAMPLE = 4


@dataclass
class ScoredCandidate:
    track_uuid: uuid.UUID
    aligned_hashes: int
    offset_seconds: float | None
    confidence: float = 0.0


def pcm_duration_sec(pcm: bytes) -> float:
    return (len(pcm) // BYTES_PER_SAMPLE) / SAMPLE_RATE


def extract_pcm_window(pcm: bytes, start_sec: float, stop_sec: float) -> bytes:
    lo = int(start_sec * SAMPLE_RATE) * BYTES_PER_SAMPLE
    hi = int(stop_sec * SAMPLE_RATE) * BYTES_PER_SAMPLE
    lo = max(0, min(lo, len(pcm)))
    hi = max(lo, min(hi, len(pcm)))
    return pcm[lo:hi]


def normalize_confidence(aligned_hashes: int) -> float:
    return 0.0 if aligned_hashes <= 0 else min(aligned_hashes / STRONG_MATCH_HASHES, 1.0)


def _group(rows_with_window):
    """(window, row) pairs -> {stripped reference_path: [(window, row)]} in first-seen order; non-UUID paths dropped later."""
    groups: dict[str, list] = {}
    for w, m in rows_with_window:
        groups.setdefault(m.reference_path.strip(), []).append((w, m))
    return groups


def consensus_score(window_results: Sequence[Sequence[OlafMatch]]) -> list[ScoredCandidate]:
    out = []
    for path, wm in _group((w, m) for w, rows in enumerate(window_results) for m in rows).items():
        try:
            tid = uuid.UUID(path)
        except ValueError:
            continue
        total = sum(m.match_count for _, m in wm)
        offset = statistics.median([m.reference_start for _, m in wm])
        if len({w for w, _ in wm}) < 2:
            total = max(total // 2, 1)
        out.append(ScoredCandidate(tid, total, offset))
    return out


def matches_to_candidates(matches: Sequence[OlafMatch]) -> list[ScoredCandidate]:
    out = []
    for path, wm in _group((0, m) for m in matches).items():
        try:
            tid = uuid.UUID(path)
        except ValueError:
            continue
        out.append(ScoredCandidate(tid, sum(m.match_count for _, m in wm), statistics.median([m.reference_start for _, m in wm])))
    return out


def plan_windows(pcm: bytes) -> tuple[bool, list[bytes]]:
    """(is_short, engine calls the reference would make for this clip, in order; b"" = call skipped)."""
    dur = pcm_duration_sec(pcm)
    if dur > SHORT_CLIP_THRESHOLD_SEC:
        return False, [pcm]
    wins = []
    for a, b in SUB_WINDOWS:
        stop = min(b, dur)
        wins.append(extract_pcm_window(pcm, a, stop) if a < stop else b"")
    return True, wins


def plan_window_ranges(pcm: bytes) -> tuple[bool, list[tuple[int, int] | None]]:
    """plan_windows in samples: (is_short, [(first sample, one-past-last sample) or None per engine call]) -- the same
    slicing arithmetic (exact.py:374-399: int(sec * 16000), clamped), without copying the bytes out."""
    dur = pcm_duration_sec(pcm)
    n = len(pcm) // BYTES_PER_SAMPLE
    if dur > SHORT_CLIP_THRESHOLD_SEC:
        return False, [(0, n)]
    out: list[tuple[int, int] | None] = []
    for a, b in SUB_WINDOWS:
        stop = min(b, dur)
        if not a < stop:
            out.append(None)
            continue
        lo = max(0, min(int(a * SAMPLE_RATE) * BYTES_PER_SAMPLE, len(pcm)))
        hi = max(lo, min(int(stop * SAMPLE_RATE) * BYTES_PER_SAMPLE, len(pcm)))
        out.append((lo // BYTES_PER_SAMPLE, hi // BYTES_PER_SAMPLE) if hi > lo else None)
    return True, out


def rank(cands: list[ScoredCandidate], max_results: int) -> list[ScoredCandidate]:
    kept = [c for c in cands if c.aligned_hashes >= MIN_ALIGNED_HASHES]
    for c in kept:
        c.confidence = normalize_confidence(c.aligned_hashes)
    kept.sort(key=lambda c: c.confidence, reverse=True)
    return kept[:max_results]


def score_clips(clips: Sequence[bytes], max_results: int = 10,
                query_many: Callable[[Sequence[bytes]], list[list[OlafMatch]]] | None = None) -> list[list[ScoredCandidate]]:
    """Exact-lane scoring for a batch of clips with one engine call for all their windows. With the default engine
    the sub-windows are passed as offsets into their clip (fingerprint.query_windows_sync), so a 5 s clip crosses PCIe
    once instead of as three overlapping 3.5 s copies."""
    if query_many is None:
        from .fingerprint import query_windows_sync
        ranges = [plan_window_ranges(c) if c else (False, []) for c in clips]
        wins = [(i, r[0], r[1]) for i, (_, rs) in enumerate(ranges) for r in rs if r is not None]
        rows_w = iter(query_windows_sync(clips, wins)) if wins else iter(())
        out_w = []
        for short, rs in ranges:
            res = [next(rows_w) if r is not None else [] for r in rs]
            out_w.append([] if not rs else rank(consensus_score(res) if short else matches_to_candidates(res[0]), max_results))
        return out_w
    plans = [plan_windows(c) if c else (False, []) for c in clips]
    flat = [w for _, wins in plans for w in wins if w]
    rows = iter(query_many(flat)) if flat else iter(())
    out = []
    for short, wins in plans:
        res = [next(rows) if w else [] for w in wins]
        if not wins:
            out.append([])
        elif short:
            out.append(rank(consensus_score(res), max_results))
        else:
            out.append(rank(matches_to_candidates(res[0]), max_results))
    return out
