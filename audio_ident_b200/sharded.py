"""Identification against an index sharded over the GPUs of one box (one process per GPU, torch.distributed).

Sharding (DESIGN.md "Multi-GPU"): tracks are dealt to ranks round-robin (global track g lives on rank g % P) and
each rank holds ordinary index segments for its own tracks. A (track, offset) vote histogram therefore lives
entirely on one rank, every rank can apply the AID_MIN_VOTES threshold and keep its exact top-50 locally, and what
has to meet is the per-rank rows; every rank then performs the same merge: order by (count desc, global track asc,
offset asc), keep 50. The result is bit-identical to a single index holding all tracks (tests/test_sharded_cpu.py,
tests/test_gpu_sharded.py, tests/test_gpu_ipc_exchange.py).

Query paths, in the order `query` prefers them:
  * fused (product path; `enable_peer_exchange`): one engine call, aid_identify_exchange_dev / _host -- split
    fingerprinting, fingerprints and rows exchanged through peer memory from inside the kernels, device merge;
  * NCCL (`_query_device`; comparison, and batches larger than the exchange was sized for): all-gather of padded
    fingerprints, local probe, all-gather of 50-row blocks, torch sort;
  * host split (`device=False`, gloo tests on CPU backends): the same exchange through host arrays.
Hash-range sharding, which BASELINE.json's north_star sketches, would split every histogram across ranks, so no
rank could threshold and all partial (track, offset, count) tuples -- not 50 rows -- would have to cross NVLink;
DESIGN.md section 5 has the byte counts and the measurement behind that decision.

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


def connect_local(identifiers) -> None:
    """Wires the peer exchanges of ranks that live in one process (each created with
    enable_peer_exchange(..., connect=False)), in rank order."""
    xs = [sh._xchg for sh in identifiers]
    for x in xs:
        x.connect_local(xs)


class ShardedIdentifier:
    """`backend` is an audio_ident_b200.engine.Engine (or anything with the same index_add / query /
    query_hashes methods, which is how the CPU gloo tests drive the exchange logic). With `device` set (a torch
    CUDA device) row blocks are mapped, exchanged (NCCL) and merged on the GPU."""

    def __init__(self, backend, rank: int = 0, world: int = 1, group=None, device=None):
        self.backend, self.rank, self.world, self.group, self.device = backend, rank, world, group, device
        self.to_global: list[int] = []          # local track number -> global track number
        self._to_global_dev = None
        self._map_dev = None                    # the same table as int32 for the fused exchange kernels
        self._stream = None
        self._xchg = None                       # engine.Exchange once enable_peer_exchange() ran

    # ---- fused rank + exchange over peer memory (include/audio_ident_b200.h, aid_exchange_*)
    def enable_peer_exchange(self, max_queries: int, connect: bool = True) -> None:
        """Replaces "all-gather the 50-row blocks with NCCL, sort them with torch" by the engine's own exchange:
        k_rank stores its rows into every rank's receive window over NVLink, k_merge_blocks merges on the device.
        One process per GPU: the 64-byte IPC handles travel once through torch.distributed; ranks living in one
        process (tests) pass connect=False and call sharded.connect_local([...]) afterwards."""
        if self.device is None or not hasattr(self.backend, "exchange"):
            raise RuntimeError("the peer exchange needs CUDA engines")
        self._xchg = self.backend.exchange(self.rank, self.world, int(max_queries))
        if connect and self.world > 1:
            import torch.distributed as dist
            handles = [None] * self.world
            dist.all_gather_object(handles, self._xchg.handle(), group=self.group)
            self._xchg.connect(handles)
            dist.barrier(group=self.group)

    def _rows_out(self, rows, nrows):
        """int32 [n, 50, 5] merged rows + counts from the engine -> the (int64 rows with -1 fill, n_rows) merge_rows returns."""
        import torch
        valid = torch.arange(MAX_ROWS, device=rows.device)[None, :] < nrows[:, None]
        r64 = rows.to(torch.int64)
        r64[:, :, 1] &= 0xFFFFFFFF
        return torch.where(valid[:, :, None], r64, torch.full_like(r64, -1)), nrows

    # ---- ingest: no collective
    def my_tracks(self, n_global: int, first: int = 0) -> list[int]:
        return [g for g in range(first, first + n_global) if shard_of(g, self.world) == self.rank]

    def _register(self, global_ids, ok):
        assert all(shard_of(g, self.world) == self.rank for g in global_ids)
        assert not self.to_global or not len(global_ids) or global_ids[0] > self.to_global[-1], "add in increasing global order"
        self.to_global.extend(int(g) for g, k in zip(global_ids, ok) if k)      # a refused track gets no local number
        self._to_global_dev = None
        self._map_dev = None

    def add(self, pcm, sample_off, global_ids: Sequence[int], device: bool = False) -> np.ndarray:
        ok = self.backend.index_add(pcm, sample_off, [str(g) for g in global_ids], device=device)
        self._register(global_ids, ok)
        return ok

    def add_hashes(self, h, t, hash_off, n_frames, global_ids: Sequence[int]) -> np.ndarray:
        ok = self.backend.index_add_hashes(h, t, hash_off, n_frames, [str(g) for g in global_ids])
        self._register(global_ids, ok)
        return ok

    # ---- identify: local probe + one all-gather + identical merge everywhere
    def _all_gather(self, t):
        import torch
        import torch.distributed as dist
        out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
        return out.reshape((self.world,) + tuple(t.shape))

    def _merge_device(self, raw, nn):
        """raw int32 [n_q, max_rows, 5] + counts on the device -> global numbering -> exchange -> merge (all on the GPU)."""
        import torch
        if self._to_global_dev is None:
            self._to_global_dev = torch.tensor(self.to_global if self.to_global else [0], dtype=torch.int64, device=self.device)
        n_q = raw.shape[0]
        blk = torch.full((n_q, MAX_ROWS, 5), -1, dtype=torch.int64, device=self.device)
        blk[:, :raw.shape[1]] = raw.to(torch.int64)
        blk[:, :, 1] = self._to_global_dev[(blk[:, :, 1] & 0xFFFFFFFF).clamp_(0, len(self._to_global_dev) - 1)]
        valid = torch.arange(MAX_ROWS, device=self.device)[None, :] < nn[:, None]
        blk = torch.where(valid[:, :, None], blk, torch.full_like(blk, -1))
        if self.world == 1:
            return merge_rows(blk[None])
        return merge_rows(self._all_gather(blk))

    def _query_fused(self, d_pcm, sample_off):
        """Whole identification step as one engine call (aid_identify_exchange_dev): split fingerprinting, fingerprints
        and rows exchanged through the ranks' windows over NVLink, merge on the device. No NCCL call, no host
        synchronisation; the only torch work is shaping the result."""
        import torch
        n = len(sample_off[0]) if isinstance(sample_off, tuple) else len(sample_off) - 1
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=self.device)
        self._stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self._stream):
            i32 = dict(dtype=torch.int32, device=self.device)
            if self._map_dev is None:
                self._map_dev = torch.tensor(self.to_global if self.to_global else [0], **i32)
            rows = torch.empty((n, MAX_ROWS, 5), **i32)
            nrows = torch.empty(n, **i32)
            self.backend.identify_exchange_dev(self._xchg, d_pcm, sample_off, self._map_dev, len(self.to_global), rows,
                                               nrows, MAX_ROWS, self._stream.cuda_stream)
            out = self._rows_out(rows, nrows)
        torch.cuda.current_stream(self.device).wait_stream(self._stream)
        return out

    def _query_device(self, d_pcm, sample_off):
        """The same step with NCCL (kept for comparison and for batches larger than the exchange was sized for): split fingerprinting, NCCL all-gather of the hashes
        (padded to the largest rank), local probe/vote on device buffers, all-gather of the row blocks, merge.
        One host synchronisation (the per-rank hash totals, needed to size the exchange buffer)."""
        import torch
        P, r = self.world, self.rank
        n = len(sample_off) - 1
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=self.device)
        self._stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self._stream):
            st = self._stream.cuda_stream
            bounds = [q * n // P for q in range(P + 1)]
            lo, hi = bounds[r], bounds[r + 1]
            per = max(bounds[q + 1] - bounds[q] for q in range(P))
            res = self.backend.fingerprint_dev(d_pcm, sample_off[lo:hi + 1], st)
            i32 = dict(dtype=torch.int32, device=self.device)
            off = torch.zeros(per + 1, **i32)
            stat = torch.zeros(per, **i32)
            self.backend.copy_device(off, res.d_hash_off, 4 * (hi - lo + 1), st)
            self.backend.copy_device(stat, res.d_status, 4 * (hi - lo), st)
            lens = off[1:] - off[:-1]
            lens = torch.where((stat & 3) != 0, torch.zeros_like(lens), lens)
            lens[hi - lo:] = 0
            meta = torch.cat([off[hi - lo:hi - lo + 1], off[:-1], lens])          # total, begins[per], lens[per]
            metas = self._all_gather(meta) if P > 1 else meta[None]
            totals = metas[:, 0].tolist()                                           # the one host synchronisation
            cap = max(int(max(totals)), 1)
            buf = torch.zeros((2, cap), **i32)
            self.backend.copy_device(buf[0], res.d_hash, 4 * totals[r], st)
            self.backend.copy_device(buf[1], res.d_t_anchor, 4 * totals[r], st)
            allb = self._all_gather(buf) if P > 1 else buf[None]                    # [P, 2, cap]
            begins = torch.cat([metas[q, 1:1 + bounds[q + 1] - bounds[q]] + q * 2 * cap for q in range(P)]).contiguous()
            lens_all = torch.cat([metas[q, 1 + per:1 + per + bounds[q + 1] - bounds[q]] for q in range(P)]).contiguous()
            rows = torch.empty((n, MAX_ROWS, 5), **i32)
            nrows = torch.empty(n, **i32)
            if self._xchg is not None and n <= self._xchg.max_queries:
                if self._map_dev is None:
                    self._map_dev = torch.tensor(self.to_global if self.to_global else [0], **i32)
                self.backend.match_exchange_dev(self._xchg, allb.data_ptr(), allb.data_ptr() + 4 * cap, begins, lens_all,
                                                None, n, self._map_dev, len(self.to_global), rows, nrows, MAX_ROWS, st)
                out = self._rows_out(rows, nrows)
            else:
                self.backend.match_dev(allb.data_ptr(), allb.data_ptr() + 4 * cap, begins, lens_all, None, n, rows, nrows,
                                       MAX_ROWS, st)
                out = self._merge_device(rows, nrows)
        torch.cuda.current_stream(self.device).wait_stream(self._stream)
        return out

    def _merge_local(self, rows: np.ndarray, n: np.ndarray):
        """this rank's structured rows -> global numbering -> exchange -> merged [n_q, 50, 5], n_rows."""
        if self.device is None:                                     # CPU path (gloo tests, single process)
            local = rows_to_array(rows, n, np.asarray(self.to_global, np.int64))
            if self.world == 1:
                return merge_rows(local[None])
            import torch
            return merge_rows(self._all_gather(torch.from_numpy(local)).numpy())
        import torch
        n_q = rows.shape[0]
        raw = torch.from_numpy(rows.view(np.int32).reshape(n_q, rows.shape[1], 5)).to(self.device, non_blocking=True)
        return self._merge_device(raw, torch.from_numpy(n).to(self.device, non_blocking=True))

    def check(self) -> None:
        """Raises EngineError if a step of the peer exchange gave up waiting for a rank (AID_E_TIMEOUT) or a rank's
        fingerprints did not fit the window (AID_E_CAPACITY). Synchronises the device; reports a failure once."""
        if self._xchg is not None:
            self._xchg.check()

    def query_host(self, pcm, sample_off, rows_first: int = 0, rows_count: int | None = None):
        """Host window PCM in, host rows out, through the fused exchange (aid_identify_exchange_host): only this
        rank's slice of the batch crosses PCIe. `pcm` is a host array (or an address such that pcm[sample_off[i]] is the
        first sample of window i for this rank's slice). Returns (rows MATCH_ROW_DTYPE [count, 50] with GLOBAL track
        numbers, n_rows int32 [count]) for windows [rows_first, rows_first + rows_count). Raises on a failed exchange."""
        import torch
        from ._lib import MATCH_ROW_DTYPE
        if self._xchg is None:
            raise RuntimeError("enable_peer_exchange() first")
        if isinstance(sample_off, tuple):                    # (win_begin, win_end): overlapping windows, one copy of the span
            sample_off = (np.ascontiguousarray(sample_off[0], np.int64), np.ascontiguousarray(sample_off[1], np.int64))
            n = len(sample_off[0])
        else:
            sample_off = np.ascontiguousarray(sample_off, np.int64)
            n = len(sample_off) - 1
        cnt = n - rows_first if rows_count is None else int(rows_count)
        if self._map_dev is None:
            self._map_dev = torch.tensor(self.to_global if self.to_global else [0], dtype=torch.int32, device=self.device)
        if getattr(self, "_host_rows", None) is None or self._host_rows.shape[0] < cnt:
            self._host_rows_t = torch.empty((max(cnt, 1), MAX_ROWS * 5), dtype=torch.int32, pin_memory=True)
            self._host_n_t = torch.empty(max(cnt, 1), dtype=torch.int32, pin_memory=True)
            self._host_rows = self._host_rows_t.numpy().view(MATCH_ROW_DTYPE).reshape(-1, MAX_ROWS)
        self.backend.identify_exchange_host(self._xchg, pcm, sample_off, self._map_dev, len(self.to_global),
                                            self._host_rows_t, self._host_n_t, rows_first, cnt)
        return self._host_rows[:cnt], self._host_n_t.numpy()[:cnt]

    def query(self, pcm, sample_off, device: bool = False, split_fingerprint: bool = True, check: bool = True):
        """Every rank passes the same query batch. With several ranks the fingerprinting itself is split (rank r
        fingerprints windows [r*n/P, (r+1)*n/P)) so that no stage is replicated. With the peer exchange enabled the
        step is asynchronous; check=True (default) waits for it and raises if a rank failed to deliver -- a failed
        exchange must not read as "no match". Callers that pipeline steps pass check=False and call check() later."""
        if isinstance(sample_off, tuple):                    # (win_begin, win_end) over device PCM: the fused path only
            if not (device and self._xchg is not None and len(sample_off[0]) <= self._xchg.max_queries):
                raise ValueError("overlapping windows need device PCM and the peer exchange (enable_peer_exchange)")
            out = self._query_fused(pcm, (np.ascontiguousarray(sample_off[0], np.int64), np.ascontiguousarray(sample_off[1], np.int64)))
            if check:
                self._xchg.check()
            return out
        sample_off = np.ascontiguousarray(sample_off, np.int64)
        n = len(sample_off) - 1
        if device and self.device is not None and split_fingerprint and n >= self.world and hasattr(self.backend, "match_dev"):
            if self._xchg is not None and n <= self._xchg.max_queries:
                out = self._query_fused(pcm, sample_off)
                if check:
                    self._xchg.check()
                return out
            return self._query_device(pcm, sample_off)
        if self.world == 1 or not split_fingerprint or n < self.world:
            rows, nr = self.backend.query(pcm, sample_off, device=device)
            return self._merge_local(rows, nr)
        import torch
        lo, hi = self.rank * n // self.world, (self.rank + 1) * n // self.world
        sub_off = sample_off[lo:hi + 1]
        if device:
            res = self.backend.fingerprint_dev(pcm, sub_off)
            off32 = self.backend.to_host(res.d_hash_off, hi - lo + 1, np.uint32).astype(np.int64)
            h = self.backend.to_host(res.d_hash, int(off32[-1]), np.uint32)
            t = self.backend.to_host(res.d_t_anchor, int(off32[-1]), np.uint32)
            st = self.backend.to_host(res.d_status, hi - lo, np.int32)
            keep = (st & 3) == 0
        else:
            h, t, off32, st = self.backend.fingerprint(pcm, sub_off)
            keep = np.ones(hi - lo, bool)
        cnt = np.where(keep, np.diff(off32), 0)
        if not keep.all():
            sel = np.concatenate([np.arange(off32[i], off32[i + 1]) for i in range(hi - lo) if keep[i]]) if keep.any() else np.zeros(0, np.int64)
            h, t = h[sel], t[sel]
        # all-gather window counts (fixed size), then the hashes padded to the largest rank
        dev = self.device if self.device is not None else "cpu"
        per = (n + self.world - 1) // self.world + 1
        c = torch.zeros(per, dtype=torch.int64, device=dev); c[:hi - lo] = torch.from_numpy(cnt).to(dev)
        counts = self._all_gather(c)                                       # [P, per]
        totals = counts.sum(dim=1)
        cap = int(totals.max().item())
        buf = torch.zeros((2, max(cap, 1)), dtype=torch.int64, device=dev)
        buf[0, :len(h)] = torch.from_numpy(h.astype(np.int64)).to(dev)
        buf[1, :len(t)] = torch.from_numpy(t.astype(np.int64)).to(dev)
        allb = self._all_gather(buf).cpu().numpy()                         # [P, 2, cap]
        counts = counts.cpu().numpy(); totals = totals.cpu().numpy()
        hs, ts, lens = [], [], []
        for r in range(self.world):
            nr_ = (r + 1) * n // self.world - r * n // self.world
            hs.append(allb[r, 0, :totals[r]]); ts.append(allb[r, 1, :totals[r]]); lens.append(counts[r, :nr_])
        off = np.concatenate([[0], np.cumsum(np.concatenate(lens))])
        rows, nr = self.backend.query_hashes(np.concatenate(hs).astype(np.uint32), np.concatenate(ts).astype(np.uint32), off)
        return self._merge_local(rows, nr)

    def query_hashes(self, h, t, hash_off):
        rows, n = self.backend.query_hashes(h, t, hash_off)
        return self._merge_local(rows, n)
