"""Long recordings (BASELINE.json configs[4]): time-axis splitting with a halo, and a continuous match stream.

The reference has no counterpart: its ingest refuses anything above 30 minutes (app/ingest/pipeline.py:41, :139-143)
and its query path simply hands the whole file to the engine ("long-context"). Two tools here:

* ``fingerprint_chunked`` -- fingerprint one long recording in time slices that can live on different GPUs, with
  results IDENTICAL to a single pass. A slice that owns anchor frames [A, B) is fingerprinted on frames
  [A - 12, B + 33 + 12): 12 frames of halo make every peak the slice uses see its full 25-frame window, and the 33
  extra frames hold every target an anchor in [A, B) can pair with. Hashes whose anchor lies outside [A, B) are
  dropped, times are shifted back to the recording's frame axis. (The 2048-peaks-per-256-frame-block capacity rule
  is evaluated on the slice's own block grid; it only matters for degenerate, tie-heavy input.)
* ``identify_stream`` -- cut the recording into overlapping vote windows, identify all of them in one sharded batch
  (fingerprinting split across ranks, index sharded by track), and stitch consecutive windows that agree on
  (track, offset) into segments.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

HOP, NFFT = 128, 1024
HALO_T, DT_MAX = 12, 33


def num_frames(n_samples: int) -> int:
    return 0 if n_samples < NFFT else (n_samples - NFFT) // HOP + 1


def plan_slices(total_frames: int, n_slices: int, align: int = 256) -> list[tuple[int, int]]:
    """Anchor-frame ranges [A, B) covering [0, total_frames), boundaries on multiples of `align`."""
    n_slices = max(1, min(n_slices, max(1, total_frames // align)))
    per = -(-total_frames // n_slices)
    per = -(-per // align) * align
    out, a = [], 0
    while a < total_frames:
        out.append((a, min(total_frames, a + per)))
        a += per
    return out


def slice_samples(a: int, b: int, total_frames: int) -> tuple[int, int, int]:
    """(first sample, one-past-last sample, first frame) of the extended slice for anchors [a, b)."""
    e0 = max(0, a - HALO_T)
    e1 = min(total_frames, b + DT_MAX + HALO_T)
    return e0 * HOP, (e1 - 1) * HOP + NFFT, e0


def fingerprint_chunked(engine, pcm: np.ndarray, n_slices: int, slices=None):
    """(hash, t_anchor) of the whole recording, computed slice by slice; equal to engine.fingerprint(pcm) as sets and,
    after the final sort by (t_anchor, position), in order."""
    pcm = np.ascontiguousarray(pcm, np.float32)
    T = num_frames(len(pcm))
    slices = slices if slices is not None else plan_slices(T, n_slices)
    clips, meta = [], []
    for a, b in slices:
        s0, s1, e0 = slice_samples(a, b, T)
        clips.append(pcm[s0:s1]); meta.append((a, b, e0))
    off = np.zeros(len(clips) + 1, np.int64); off[1:] = np.cumsum([len(c) for c in clips])
    h, t, hoff, st = engine.fingerprint(np.concatenate(clips) if clips else pcm[:0], off)
    hs, ts = [], []
    for i, (a, b, e0) in enumerate(meta):
        if st[i] & 3:
            raise RuntimeError(f"slice {i} failed with status {int(st[i])}")
        hh, tt = h[hoff[i]:hoff[i + 1]], t[hoff[i]:hoff[i + 1]].astype(np.int64) + e0
        keep = (tt >= a) & (tt < b)
        hs.append(hh[keep]); ts.append(tt[keep].astype(np.uint32))
    return (np.concatenate(hs) if hs else h[:0]), (np.concatenate(ts) if ts else t[:0])


@dataclass
class Segment:
    track: int            # global track number
    offset_frames: int    # t_ref - t_recording
    start_s: float        # where in the recording the match starts / ends
    stop_s: float
    votes: int            # aligned hashes summed over the windows of the segment
    windows: int


def plan_windows(n_samples: int, window_s: float = 10.0, hop_s: float = 5.0, sr: int = 16000) -> np.ndarray:
    """Start samples of the vote windows (hop-aligned so window frames coincide with recording frames)."""
    w, h = int(window_s * sr), int(hop_s * sr) // HOP * HOP
    if n_samples <= w:
        return np.zeros(1, np.int64)
    return np.arange(0, n_samples - w + h, h, dtype=np.int64).clip(max=max(0, n_samples - w) // HOP * HOP)


def identify_stream(identifier, pcm, n_samples: int, window_s: float = 10.0, hop_s: float = 5.0, device: bool = False,
                    min_votes: int = 8, offset_tol: int = 2):
    """`identifier` is a sharded.ShardedIdentifier; `pcm` a host array or a device pointer (device=True).
    Returns (segments, merged_rows, n_rows, window_starts)."""
    starts = np.unique(plan_windows(n_samples, window_s, hop_s))
    w = min(int(window_s * 16000), n_samples)
    # windows overlap, so they are passed as a ragged batch that re-uses the same buffer: query() takes offsets
    # only for dense batches, hence one gather into a window-major buffer
    if device:
        import torch
        src = pcm if hasattr(pcm, "data_ptr") else None
        assert src is not None, "device=True expects a torch CUDA tensor"
        if getattr(identifier, "_xchg", None) is not None and len(starts) <= identifier._xchg.max_queries:
            # the windows are fingerprinted where they lie in the recording (overlapping ranges, no gather, no copy)
            merged, n = identifier.query(src.data_ptr(), (starts, starts + w), device=True)
        else:
            idx = torch.from_numpy(starts).to(src.device)[:, None] + torch.arange(w, device=src.device)[None, :]
            buf = src[idx].contiguous()
            off = np.arange(len(starts) + 1, dtype=np.int64) * w
            merged, n = identifier.query(buf.data_ptr(), off, device=True)
        m = merged.cpu().numpy(); nn = n.cpu().numpy()
    else:
        pcm = np.ascontiguousarray(pcm, np.float32)
        buf = np.concatenate([pcm[s:s + w] for s in starts])
        off = np.arange(len(starts) + 1, dtype=np.int64) * w
        merged, n = identifier.query(buf, off)
        m = np.asarray(merged.cpu().numpy() if hasattr(merged, "cpu") else merged); nn = np.asarray(n.cpu().numpy() if hasattr(n, "cpu") else n)
    return _stitch(m, nn, starts, w, min_votes, offset_tol), m, nn, starts


def _stitch(m: np.ndarray, nn: np.ndarray, starts: np.ndarray, w: int, min_votes: int, offset_tol: int) -> list[Segment]:
    """Consecutive windows whose best row names the same track at (nearly) the same recording offset form one segment.
    Definition (the loop this vectorises): a window continues the current segment iff its best row has >= min_votes,
    the same track, and an offset within offset_tol frames of the segment's FIRST window."""
    n = len(starts)
    if n == 0:
        return []
    valid = (nn > 0) & (m[:, 0, 0] >= min_votes)
    track = m[:, 0, 1]
    off_rec = m[:, 0, 2] - starts // HOP                       # offset relative to the recording's frame axis
    votes = m[:, 0, 0]
    # candidate runs: maximal stretches of valid windows with one track (offsets are checked per run below)
    same = valid[1:] & valid[:-1] & (track[1:] == track[:-1])
    run_start = np.flatnonzero(valid & np.concatenate([[True], ~same]))
    run_end = np.flatnonzero(valid & np.concatenate([~same, [True]])) + 1
    segs: list[Segment] = []
    for a, b in zip(run_start, run_end):
        k = int(a)
        while k < b:                                           # split a run where the offset leaves the tolerance
            drift = np.abs(off_rec[k:b] - off_rec[k]) > offset_tol
            e = k + (int(np.argmax(drift)) if drift.any() else int(b - k))
            segs.append(Segment(int(track[k]), int(off_rec[k]), float(starts[k]) / 16000.0, float(starts[e - 1] + w) / 16000.0,
                                int(votes[k:e].sum()), e - k))
            k = e
    return segs
