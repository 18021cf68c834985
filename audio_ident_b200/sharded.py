"""Identification against an index sharded over the GPUs of one box (one process per GPU, torch.distributed).

Sharding (DESIGN.md "Multi-GPU"): tracks are dealt to ranks round-robin (global track g lives on rank g % P) and
each rank holds ordinary index segments for its own tracks. A (track, offset) vote histogram therefore lives
entirely on one rank, every rank can apply the AID_MIN_VOTES threshold and keep its exact top-50 locally, and the
only exchange is one all-gather of the fixed-size per-rank row blocks (50 rows x 20 B per query) over
NCCL/NVLink, after which every rank performs the same merge: order by (count desc, global track asc, offset asc),
keep 50. The result is bit-identical to a single index holding all tracks (tests/test_sharded_cpu.py,
tests/test_gpu_sharded.py). Hash-range sharding, which BASELINE.json's north_star sketches, would split every
histogram across ranks, so no rank could threshold and all partial (track, offset, count) tuples -- not 50 rows --
would have to cross NVLink; see DESIGN.md for the byte counts behind that decision.

Bulk ingest needs no collective at all: each rank fingerprints and stores its own tracks.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

MAX_ROWS = 50
ROW_FIELDS = ("count", "track", "offset", "q_first", "q_last")


def shard_of(global_track: int, world: int) -> int:
    return global_track % world


def rows_to_array(rows: np.ndarray, n: np.ndarray, to_global: np.ndarray) -> np.ndarray:
    """structured rows [n_q, 50] + counts -> int64 [n_q, 50, 5] with global track numbers; unused rows count = -1."""
    n_q = rows.shape[0]
    out = np.full((n_q, MAX_ROWS, 5), -1, np.int64)
    for j, name in enumerate(ROW_FIELDS):
        out[:, :rows.shape[1], j] = rows[name]
    if len(to_global):
        tr = np.clip(out[:, :, 1], 0, len(to_global) - 1)
        out[:, :, 1] = to_global[tr]
    valid = np.arange(MAX_ROWS)[None, :] < n[:, None]
    out[~valid] = -1
    return out


_OFF_BIAS = 1 << 18          # offsets are > -2^18


def _sort_keys(r):
    """(count desc, track asc, offset asc) as one ascending int64 key; rows with count < 0 sort last.
    Works on numpy arrays and torch tensors alike (r[..., 5] integer)."""
    count, track, offset = r[..., 0], r[..., 1], r[..., 2]
    key = ((0xFFFFF - count.clip(0, 0xFFFFF)) << 44) | (track.clip(0, (1 << 25) - 1) << 19) | (offset + _OFF_BIAS).clip(0, (1 << 19) - 1)
    return key + (count < 0) * (1 << 62)


def merge_rows(blocks):
    """blocks int64 [P, n_q, 50, 5] (numpy or torch) -> (merged [n_q, 50, 5], n_rows [n_q]); the same arithmetic on
    every rank, vectorised over the whole query batch."""
    if isinstance(blocks, np.ndarray):
        P, n_q = blocks.shape[0], blocks.shape[1]
        allr = np.transpose(blocks, (1, 0, 2, 3)).reshape(n_q, P * MAX_ROWS, 5)
        order = np.argsort(_sort_keys(allr), axis=1, kind="stable")[:, :MAX_ROWS]
        merged = np.take_along_axis(allr, order[:, :, None], axis=1)
        n_rows = (merged[:, :, 0] >= 0).sum(axis=1).astype(np.int32)
        merged[merged[:, :, 0] < 0] = -1
        return merged, n_rows
    import torch
    P, n_q = blocks.shape[0], blocks.shape[1]
    allr = blocks.permute(1, 0, 2, 3).reshape(n_q, P * MAX_ROWS, 5)
    order = torch.sort(_sort_keys(allr), dim=1, stable=True).indices[:, :MAX_ROWS]
    merged = torch.gather(allr, 1, order[:, :, None].expand(-1, -1, 5))
    n_rows = (merged[:, :, 0] >= 0).sum(dim=1).to(torch.int32)
    merged = torch.where(merged[:, :, :1] < 0, torch.full_like(merged, -1), merged)
    return merged, n_rows


class ShardedIdentifier:
    """`backend` is an audio_ident_b200.engine.Engine (or anything with the same index_add / query /
    query_hashes methods, which is how the CPU gloo tests drive the exchange logic)."""

    def __init__(self, backend, rank: int = 0, world: int = 1, group=None, device=None):
        self.backend, self.rank, self.world, self.group, self.device = backend, rank, world, group, device
        self.to_global: list[int] = []          # local track number -> global track number

    # ---- ingest: no collective
    def my_tracks(self, n_global: int, first: int = 0) -> list[int]:
        return [g for g in range(first, first + n_global) if shard_of(g, self.world) == self.rank]

    def add(self, pcm, sample_off, global_ids: Sequence[int], device: bool = False) -> np.ndarray:
        assert all(shard_of(g, self.world) == self.rank for g in global_ids)
        assert not self.to_global or not len(global_ids) or global_ids[0] > self.to_global[-1], "add in increasing global order"
        ok = self.backend.index_add(pcm, sample_off, [str(g) for g in global_ids], device=device)
        self.to_global.extend(int(g) for g, k in zip(global_ids, ok) if k)      # a refused track gets no local number
        return ok

    def add_hashes(self, h, t, hash_off, n_frames, global_ids: Sequence[int]) -> np.ndarray:
        ok = self.backend.index_add_hashes(h, t, hash_off, n_frames, [str(g) for g in global_ids])
        self.to_global.extend(int(g) for g, k in zip(global_ids, ok) if k)
        return ok

    # ---- identify: local probe + one all-gather + identical merge everywhere
    def _exchange(self, local: np.ndarray):
        """local rows -> [P, n_q, 50, 5]; on a CUDA device the block stays a torch tensor so the merge runs there."""
        if self.world == 1 and self.device is None:
            return local[None]
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(np.ascontiguousarray(local))
        if self.device is not None:
            t = t.to(self.device, non_blocking=True)
        if self.world == 1:
            return t[None]
        out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t, group=self.group)
        out = out.reshape((self.world,) + tuple(t.shape))
        return out if self.device is not None else out.numpy()

    def query(self, pcm, sample_off, device: bool = False):
        rows, n = self.backend.query(pcm, sample_off, device=device)
        return merge_rows(self._exchange(rows_to_array(rows, n, np.asarray(self.to_global, np.int64))))

    def query_hashes(self, h, t, hash_off):
        rows, n = self.backend.query_hashes(h, t, hash_off)
        return merge_rows(self._exchange(rows_to_array(rows, n, np.asarray(self.to_global, np.int64))))
