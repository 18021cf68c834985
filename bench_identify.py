#!/usr/bin/env python
"""bench_identify.py -- BASELINE.json configs[2] and [3]: identify 5 s noisy query excerpts against an in-HBM index
of synthetic 30 s tracks; with N ranks (torchrun) the index is sharded over the ranks (track g lives on rank g % N)
and fingerprints and rows travel between the GPUs through peer memory from inside the kernels
(audio_ident_b200/csrc/exchange.cu; audio_ident_b200/sharded.py).

Library for bench.py (the driver-run line carries an `identify` block built by `identify_block`) and a CLI:

  python bench_identify.py --tracks 100000 --queries 4096                      # configs[2], 1 GPU
  torchrun --nproc-per-node 8 ... bench_identify.py --gpus 8 --tracks 1000000  # configs[3]

Queries follow SURVEY.md section 8(d): a 5.0 s excerpt at a random *sample* offset of a random indexed track plus
white Gaussian noise at 20 dB SNR, issued the way the reference issues it (three 3.5 s windows, exact.py:48-52,
consensus = sum of aligned hashes over the windows). A query = 3 windows; strong scaling (the index size is fixed
as N grows).
  * queries_per_s / ms_per_step   window PCM resident in HBM -> merged rows on every rank, CUDA events on the
                                  launching stream, max over ranks
  * e2e                           the same step through aid_identify_exchange_host: window PCM in PINNED HOST memory
                                  (each rank copies only the slice it fingerprints), merged rows of that slice back
                                  in host memory; wall clock around the synchronous calls, max over ranks
  * k_match                       achieved GB/s on SURVEY 8(d)'s numerator (32 B directory entry or 8 B of bucket
                                  table per hash and segment + 4 B per posting touched), counted live by the kernel
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR = 16000
WINDOWS = ((0, 56000), (12000, 68000), (24000, 80000))        # exact.py:48-52 in samples
WIN = 56000


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class Ctx:
    """What every leg needs: engine, rank layout, torch device, barrier."""

    def __init__(self, eng, rank, world, dev, seed=42, seconds=30.0, snr_db=20.0):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.eng, self.rank, self.world, self.dev = eng, rank, world, dev
        self.seed, self.seconds, self.snr_db = seed, seconds, snr_db
        self.samples = int(seconds * SR)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]


def build_shard(cx: Ctx, sh, tracks: int, ingest_chunk: int = 512) -> float:
    """This rank's shard of a `tracks`-track index: generate on the device, fingerprint, append (no collective)."""
    torch, eng = cx.torch, cx.eng
    t0 = time.perf_counter()
    mine = np.arange(cx.rank, tracks, cx.world, dtype=np.int64)
    buf = torch.empty(ingest_chunk * cx.samples, dtype=torch.float32, device=cx.dev)
    off_full = np.arange(ingest_chunk + 1, dtype=np.int64) * cx.samples
    for c0 in range(0, len(mine), ingest_chunk):
        ids = mine[c0:c0 + ingest_chunk]
        eng.synth_tracks(buf.data_ptr(), int(ids[0]), len(ids), cx.samples, cx.seed, stride=cx.world)   # g = rank + world * j
        ok = sh.add(buf.data_ptr(), off_full[:len(ids) + 1], [int(g) for g in ids], device=True)
        assert ok.all()
    eng.index_commit()
    cx.barrier()
    del buf
    return time.perf_counter() - t0


def make_queries(cx: Ctx, tracks: int, n_queries: int, salt: int, keep_clips: bool = False):
    """n_queries noisy 5 s excerpts (identical on every rank: same seed) ->
    (windows tensor [Q, 3, 56000] on the device, true track, true start sample, the clips [Q, 80000] if keep_clips)."""
    torch, eng, dev = cx.torch, cx.eng, cx.dev
    rng = np.random.default_rng(cx.seed + 10**6 + salt)
    q_track = rng.integers(0, tracks, n_queries)
    q_start = rng.integers(0, cx.samples - 80000 + 1, n_queries)
    gen = torch.Generator(device=dev); gen.manual_seed(cx.seed + 7 + salt)
    wins = torch.empty((n_queries, 3, WIN), dtype=torch.float32, device=dev)
    all_clips = torch.empty((n_queries, 80000), dtype=torch.float32, device=dev) if keep_clips else None
    for c0 in range(0, n_queries, 2048):                        # bounded scratch: 2048 whole tracks at a time
        c1 = min(c0 + 2048, n_queries)
        torch.cuda.synchronize()          # the engine writes `src` on its own stream: torch must be done with the last one
        src = torch.empty((c1 - c0) * cx.samples, dtype=torch.float32, device=dev)
        for j in range(c0, c1):
            eng.synth_tracks(src.data_ptr() + (j - c0) * cx.samples * 4, int(q_track[j]), 1, cx.samples, cx.seed)
        eng.sync()
        idx = torch.from_numpy(q_start[c0:c1]).to(dev)[:, None] + torch.arange(80000, device=dev)[None, :]
        clips = src.view(c1 - c0, cx.samples).gather(1, idx)
        del src, idx
        p_sig = clips.pow(2).mean(dim=1, keepdim=True)
        noise = torch.randn(clips.shape, generator=gen, device=dev) * torch.sqrt(p_sig / (10 ** (cx.snr_db / 10)))
        clips = (clips + noise).clamp_(-1.0, 1.0)
        del noise
        for w, (a_, b_) in enumerate(WINDOWS):
            wins[c0:c1, w] = clips[:, a_:b_]
        if keep_clips:
            all_clips[c0:c1] = clips
        del clips
    torch.cuda.synchronize()
    return wins, q_track, q_start, all_clips


def score(m: np.ndarray, q_track, q_start, n_queries: int):
    """Accuracy in the reference's terms (exact.py:220-293): sum aligned hashes per track over the three windows,
    top-1 with at least 8; plus a checksum over all rows. m: int64 [3Q, 50, 5], unused rows -1."""
    top1 = offs_ok = 0
    for q in range(n_queries):
        votes, first_off = {}, {}
        for w in range(3):
            r = m[3 * q + w]
            r = r[r[:, 0] >= 0]
            for cnt, tr, of in zip(r[:, 0], r[:, 1], r[:, 2]):
                votes[int(tr)] = votes.get(int(tr), 0) + int(cnt)
                first_off.setdefault((int(tr), w), int(of))
        if votes:
            best = max(votes, key=lambda k: (votes[k], -k))
            if best == int(q_track[q]) and votes[best] >= 8:
                top1 += 1
                of0 = first_off.get((best, 0))
                offs_ok += of0 is not None and abs(of0 - q_start[q] / 128.0) <= 1.0
    row_sums = (m.astype(np.int64) * np.arange(1, 6)).sum(axis=2)
    digest = int(np.bitwise_xor.reduce(row_sums.reshape(-1) & 0xFFFFFFFF))
    return top1 / n_queries, offs_ok / max(top1, 1), digest, row_sums


def rows_struct_to_array(rows: np.ndarray, n: np.ndarray) -> np.ndarray:
    """host rows (MATCH_ROW_DTYPE [k, 50]) + counts -> int64 [k, 50, 5] with -1 fill (the layout `score` reads)."""
    out = np.full((rows.shape[0], 50, 5), -1, np.int64)
    for j, name in enumerate(("count", "track", "offset", "q_first", "q_last")):
        out[:, :, j] = rows[name]
    out[np.arange(50)[None, :] >= n[:, None]] = -1
    return out


def measure(cx: Ctx, sh, tracks: int, n_queries: int, salt: int, steps: int, warmup: int, n_seg_probe_bytes: int,
            e2e: bool = True, dump: str = ""):
    """One batch size on the index `sh` holds. Returns the result dict (identical numbers on every rank)."""
    torch, eng = cx.torch, cx.eng
    wins, q_track, q_start, clips = make_queries(cx, tracks, n_queries, salt, keep_clips=e2e)
    n_win = n_queries * 3
    off = np.arange(n_win + 1, dtype=np.int64) * WIN

    def step():
        return sh.query(wins.data_ptr(), off, device=True, check=False)

    for _ in range(warmup):
        merged, n = step()
    cx.barrier()
    eng.stage_times(); eng.set_stage_timing(True); eng.match_stats()
    launches0 = eng.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()                                   # the step's own stream waits for / is waited on by this one
    for _ in range(steps):
        merged, n = step()
    ev1.record()
    torch.cuda.synchronize()
    dt_wall = time.perf_counter() - t0
    dt_dev = ev0.elapsed_time(ev1) * 1e-3
    sh.check()                                     # a rank that failed to deliver fails the bench, not the accuracy
    stage = eng.stage_times(); eng.set_stage_timing(False)
    n_hash, n_post = eng.match_stats()
    launches = eng.launches - launches0
    dt, dt_wall = cx.max_over_ranks(dt_dev, dt_wall)
    dt /= steps; dt_wall /= steps

    m = merged.cpu().numpy() if hasattr(merged, "cpu") else merged
    top1, offs_ok, digest, row_sums = score(m, q_track, q_start, n_queries)
    if dump and cx.rank == 0:                      # per-window digests and the first rows, to compare runs offline
        np.savez_compressed(f"{dump}_{n_queries}.npz", window_digest=np.bitwise_xor.reduce(row_sums & 0xFFFFFFFF, axis=1),
                            n_rows=(m[:, :, 0] >= 0).sum(axis=1), first_rows=m[:, :4, :].astype(np.int32))
    per_step = {k: v[0] / steps for k, v in stage.items()}
    res = {"queries": n_queries, "windows_per_step": n_win, "queries_per_s": n_queries / dt, "ms_per_step": dt * 1e3,
           "ms_per_step_wall": dt_wall * 1e3, "top1_accuracy": top1, "top1_offset_within_1_frame": offs_ok,
           "rows_digest": digest, "stage_ms_per_step": {k: round(v, 3) for k, v in per_step.items()},
           "gpu_launches": int(launches)}
    # ---- matcher roofline: SURVEY 8(d)'s numerator, counted by the kernel itself during the timed steps (this rank)
    t_match = stage["match"][0] / 1e3
    if t_match > 0 and stage["match"][1]:
        alg = n_hash * n_seg_probe_bytes + n_post * 4
        res["k_match"] = {"bound": "hbm (random access)", "hashes_probed_per_step": n_hash // steps,
                          "postings_touched_per_step": n_post // steps, "algorithmic_bytes_per_step": alg // steps,
                          "bytes_per_probe": n_seg_probe_bytes, "avg_launch_ms": stage["match"][0] / stage["match"][1],
                          "achieved_gbs": alg / t_match / 1e9, "rank": cx.rank}

    # ---- end to end: pinned host window PCM in (this rank's slice), merged rows of the slice back on the host
    if e2e:
        lo, hi = cx.rank * n_win // cx.world, (cx.rank + 1) * n_win // cx.world
        h_pcm = torch.empty((hi - lo) * WIN, dtype=torch.float32, pin_memory=True)
        h_pcm.copy_(wins.view(-1)[lo * WIN:hi * WIN])
        torch.cuda.synchronize()
        base = h_pcm.data_ptr() - lo * WIN * 4            # so that base[sample_off[lo]] is the slice's first sample
        for _ in range(2):
            rows_h, n_h = sh.query_host(base, off, lo, hi - lo)
        cx.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            rows_h, n_h = sh.query_host(base, off, lo, hi - lo)
        dt_e = time.perf_counter() - t0
        (dt_e,) = cx.max_over_ranks(dt_e)
        dt_e /= steps
        mine = rows_struct_to_array(rows_h, n_h)
        same = bool(np.array_equal(mine, m[lo:hi]))
        (all_same,) = cx.max_over_ranks(0.0 if same else 1.0)
        res["e2e"] = {"value": n_queries / dt_e, "unit": "queries/s", "ms_per_step": dt_e * 1e3,
                      "h2d_bytes_per_step": int(n_win * WIN * 4), "d2h_bytes_per_step": int(n_win * (50 * 20 + 4)),
                      "h2d_bytes_per_step_per_rank": int((hi - lo) * WIN * 4),
                      "rows_equal_device_path": all_same == 0.0}
        del h_pcm
        # the same queries as the batched exact lane hands them over (exact_lane.score_clips): every 5 s clip once plus
        # the three windows as offsets into it -- 2.1x less PCIe traffic for identical rows
        q_lo, q_hi = lo // 3, (hi + 2) // 3
        h_clips = torch.empty((q_hi - q_lo) * 80000, dtype=torch.float32, pin_memory=True)
        h_clips.copy_(clips.view(-1)[q_lo * 80000:q_hi * 80000])
        torch.cuda.synchronize()
        wb = (np.arange(n_queries, dtype=np.int64)[:, None] * 80000 + np.array([w[0] for w in WINDOWS], np.int64)[None, :]).reshape(-1)
        we = wb + WIN
        base_c = h_clips.data_ptr() - q_lo * 80000 * 4
        for _ in range(2):
            rows_c, n_c = sh.query_host(base_c, (wb, we), lo, hi - lo)
        cx.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            rows_c, n_c = sh.query_host(base_c, (wb, we), lo, hi - lo)
        dt_c = time.perf_counter() - t0
        (dt_c,) = cx.max_over_ranks(dt_c)
        dt_c /= steps
        same_c = bool(np.array_equal(rows_struct_to_array(rows_c, n_c), m[lo:hi]))
        (all_same_c,) = cx.max_over_ranks(0.0 if same_c else 1.0)
        res["e2e_clips"] = {"value": n_queries / dt_c, "unit": "queries/s", "ms_per_step": dt_c * 1e3,
                            "h2d_bytes_per_step": int(n_queries * 80000 * 4), "d2h_bytes_per_step": int(n_win * (50 * 20 + 4)),
                            "rows_equal_device_path": all_same_c == 0.0,
                            "note": "each 5 s clip uploaded once, its three windows passed as offsets (aid_identify_exchange_windows_host)"}
        del h_clips, clips
    del wins
    return res


def cpu_leg(cx: Ctx, cpu_tracks: int = 2048, n_queries: int = 64):
    """CPU port beside it (rank 0, N = 1): the oracle fingerprints the windows and votes against a BOUNDED index of the
    first `cpu_tracks` tracks, all host threads (one window per thread); the GPU answers the same windows against the
    same bounded index and the rows must be bit-identical (parity, asserted)."""
    from concurrent.futures import ThreadPoolExecutor
    from audio_ident_b200 import sharded
    from oracle import oracle
    torch, eng = cx.torch, cx.eng
    eng.index_clear()
    sh = sharded.ShardedIdentifier(eng, 0, 1, device=cx.dev)
    build_shard(cx, sh, cpu_tracks)
    # the bounded index on the host: fingerprints of the same tracks from the GPU (oracle.Index defines the order)
    hs, ts, trk = [], [], []
    buf = torch.empty(256 * cx.samples, dtype=torch.float32, device=cx.dev)
    off = np.arange(257, dtype=np.int64) * cx.samples
    for c0 in range(0, cpu_tracks, 256):
        c = min(256, cpu_tracks - c0)
        eng.synth_tracks(buf.data_ptr(), c0, c, cx.samples, cx.seed)
        res = eng.fingerprint_dev(buf.data_ptr(), off[:c + 1])
        hoff = eng.to_host(res.d_hash_off, c + 1, np.uint32).astype(np.int64)
        hs.append(eng.to_host(res.d_hash, int(hoff[-1]), np.uint32))
        ts.append(eng.to_host(res.d_t_anchor, int(hoff[-1]), np.uint32))
        trk.append(np.repeat(np.arange(c0, c0 + c, dtype=np.uint32), np.diff(hoff)))
    del buf
    ix = oracle.Index(np.concatenate(hs), np.concatenate(trk), np.concatenate(ts))
    wins, q_track, q_start, _ = make_queries(cx, cpu_tracks, n_queries, salt=99)
    n_win = 3 * n_queries
    off = np.arange(n_win + 1, dtype=np.int64) * WIN
    merged, n = sh.query(wins.data_ptr(), off, device=True)
    g = merged.cpu().numpy()
    pcm = wins.cpu().numpy().reshape(-1)
    threads = len(os.sched_getaffinity(0))

    def one(w):
        h, t = oracle.fingerprint(pcm[w * WIN:(w + 1) * WIN])
        return ix.match(h, t), h, t

    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(one, range(min(n_win, threads))))          # warm the tables
        t0 = time.perf_counter()
        out = list(ex.map(one, range(n_win)))
        dt = time.perf_counter() - t0
    # parity on the rows: the GPU fingerprints each window itself, so compare through the GPU's own hashes where the
    # two fingerprints agree, and the rows of every such window must be bit-identical
    res = eng.fingerprint_dev(wins.data_ptr(), off)
    hoff = eng.to_host(res.d_hash_off, n_win + 1, np.uint32).astype(np.int64)
    gh = eng.to_host(res.d_hash, int(hoff[-1]), np.uint32); gt = eng.to_host(res.d_t_anchor, int(hoff[-1]), np.uint32)
    same_fp = rows_equal = 0
    for w in range(n_win):
        rows_o, h, t = out[w]
        if np.array_equal(h, gh[hoff[w]:hoff[w + 1]]) and np.array_equal(t, gt[hoff[w]:hoff[w + 1]]):
            same_fp += 1
        else:                                   # near-tie peak flipped: vote with the GPU's fingerprints instead
            rows_o = ix.match(gh[hoff[w]:hoff[w + 1]], gt[hoff[w]:hoff[w + 1]])
        k = len(rows_o)
        ok = (g[w, :, 0] >= 0).sum() == k
        for j, name in enumerate(("count", "track", "offset", "q_first", "q_last")):
            ok = ok and np.array_equal(g[w, :k, j], rows_o[name].astype(np.int64))
        rows_equal += bool(ok)
    top1, _, _, _ = score(g, q_track, q_start, n_queries)
    del wins
    eng.index_clear()
    return {"value": n_queries / dt, "unit": "queries/s", "cores": threads, "kind": "port",
            "sample": f"{n_queries} queries (3 x 3.5 s windows each) against a bounded index of the first {cpu_tracks} "
                      f"tracks ({len(ix.hash)} postings): oracle fingerprint + aid_oracle_match, one window per thread, "
                      f"{dt:.2f} s of wall time",
            "parity": {"windows": n_win, "windows_with_bit_identical_fingerprints": same_fp,
                       "windows_with_bit_identical_rows": rows_equal, "gpu_top1_accuracy": top1}}


def identify_block(cx: Ctx, steps: int, warmup: int, index_tracks, sizes=(4096, 16384), cpu: bool = True,
                   ingest_chunk: int = 512, dump: str = ""):
    """The `identify` block of bench.py's line: for every index size, build the (sharded) index and measure every batch
    size. Returns the dict on every rank (rank 0 prints)."""
    from audio_ident_b200 import sharded
    eng = cx.eng
    out = {"metric": "queries/sec vs N-track index (a query = three 3.5 s windows of a 5 s excerpt at 20 dB SNR)",
           "unit": "queries/s", "scaling": "strong", "n_gpus": cx.world, "steps": steps, "warmup": warmup,
           "index_sharding": f"track g on rank g % {cx.world} (DESIGN.md section 5: why not hash ranges)",
           "timed_region": "device: window PCM in HBM -> fingerprint (split over ranks) -> fingerprints and rows exchanged "
                           "through peer memory -> merged rows on every rank; e2e: the same from pinned host PCM to host rows",
           "indexes": []}
    if cpu and cx.rank == 0 and cx.world == 1:
        out["cpu_baseline"] = cpu_leg(cx)
        if out["cpu_baseline"]["parity"]["windows_with_bit_identical_rows"] != out["cpu_baseline"]["parity"]["windows"]:
            raise SystemExit(f"[bench] identify parity FAILED: {out['cpu_baseline']['parity']}")
    for tracks in index_tracks:
        eng.index_clear()
        sh = sharded.ShardedIdentifier(eng, cx.rank, cx.world, device=cx.dev)
        t_build = build_shard(cx, sh, tracks, ingest_chunk)
        stats = eng.index_stats()
        log(f"[identify] rank {cx.rank}: shard {stats} built in {t_build:.1f} s")
        sh.enable_peer_exchange(3 * max(sizes))
        # bytes read per (hash, segment) probe: a 32 B directory entry when the segment is grouped, else two table words
        probe = 32 if stats["segments_grouped"] * 2 > stats["segments"] else 8
        runs = [measure(cx, sh, tracks, nq, salt, steps, warmup, probe, e2e=(salt == 0), dump=dump)
                for salt, nq in enumerate(sizes)]
        out["indexes"].append({"tracks": tracks, "tracks_per_rank": stats["tracks"], "postings_per_rank": stats["postings"],
                               "segments_per_rank": stats["segments"], "device_bytes_per_rank": stats["device_bytes"],
                               "build_seconds": t_build, "runs": runs})
        cx.barrier()                            # nobody unmaps a window a peer may still store into
        sh._xchg.close(); sh._xchg = None
        cx.barrier()
    eng.index_clear()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--tracks", type=int, default=100000, help="tracks in the whole index")
    ap.add_argument("--seconds", type=float, default=30.0)
    ap.add_argument("--queries", type=int, default=4096, help="queries per step (3 windows each)")
    ap.add_argument("--snr-db", type=float, default=20.0)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--ingest-chunk", type=int, default=512)
    ap.add_argument("--dump", default="", help="prefix of .npz files with per-window digests (debugging aid)")
    ap.add_argument("--also-queries", default="", help="comma-separated further batch sizes measured on the same index")
    ap.add_argument("--cpu", action="store_true", help="CPU leg + row parity against a bounded index (1 GPU only)")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from audio_ident_b200.engine import Engine

    rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/aid_nccl_%h_%p.log")     # NCCL's banner goes to a file, not stdout
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    eng = Engine(local_rank)
    cx = Ctx(eng, rank, world, dev, args.seed, args.seconds, args.snr_db)
    sizes = [args.queries] + [int(x) for x in args.also_queries.split(",") if x.strip()]
    blk = identify_block(cx, args.steps, args.warmup, [args.tracks], sizes, cpu=args.cpu, ingest_chunk=args.ingest_chunk,
                         dump=args.dump)
    if rank == 0:
        ix = blk["indexes"][0]; r0 = ix["runs"][0]
        print(json.dumps({
            "metric": f"queries/sec vs {args.tracks}-track index", "value": r0["queries_per_s"], "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r0["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": f"identify {args.queries} x 5 s queries (3 x 3.5 s windows, {args.snr_db:g} dB SNR) "
                                   f"against {args.tracks} x {args.seconds:g} s tracks"},
            "e2e": r0.get("e2e"), "gpu_launches": r0["gpu_launches"], "identify": blk}), flush=True)
    if world > 1:
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
