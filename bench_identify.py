#!/usr/bin/env python
"""bench_identify.py -- BASELINE.json configs[2] and [3]: identify 5 s noisy query excerpts against an in-HBM index
of synthetic 30 s tracks; with --gpus N (torchrun) the index is sharded over the ranks (track g lives on rank g % N)
and the per-rank row blocks are merged after one NCCL all-gather (audio_ident_b200/sharded.py).

  python bench_identify.py --tracks 100000 --queries 4096                      # configs[2], 1 GPU
  torchrun --nproc-per-node 8 ... bench_identify.py --gpus 8 --tracks 1000000  # configs[3]

Queries follow SURVEY.md section 8(d): a 5.0 s excerpt at a random *sample* offset of a random indexed track plus
white Gaussian noise at 20 dB SNR, issued the way the reference issues it (three 3.5 s windows, exact.py:48-52,
consensus = sum of aligned hashes over the windows). One JSON line on stdout (rank 0):
value = queries/s (a query = 3 windows) for the whole job, strong scaling (the index size is fixed as N grows).
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
WINDOWS = ((0, 56000), (12000, 68000), (24000, 80000))


def log(*a):
    print(*a, file=sys.stderr, flush=True)


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
    ap.add_argument("--exchange", choices=("peer", "nccl"), default="peer",
                    help="peer: k_rank stores rows into every rank's window over NVLink + device merge (aid_match_exchange_dev); "
                         "nccl: all-gather of 50-row blocks + torch sort (the earlier path, kept for comparison)")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from audio_ident_b200 import sharded
    from audio_ident_b200.engine import Engine

    rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.pop("NCCL_DEBUG", None)          # NCCL prints its version banner on stdout at any debug level
        if os.environ.get("AID_NCCL_DEBUG"):
            os.environ["NCCL_DEBUG"] = os.environ["AID_NCCL_DEBUG"]
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    eng = Engine(local_rank)
    sh = sharded.ShardedIdentifier(eng, rank, world, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    samples = int(args.seconds * SR)
    # ---- build this rank's shard: generate on the device, fingerprint, append (no collective)
    t0 = time.perf_counter()
    mine = np.arange(rank, args.tracks, world, dtype=np.int64)
    buf = torch.empty(args.ingest_chunk * samples, dtype=torch.float32, device=dev)
    off_full = np.arange(args.ingest_chunk + 1, dtype=np.int64) * samples
    for c0 in range(0, len(mine), args.ingest_chunk):
        ids = mine[c0:c0 + args.ingest_chunk]
        # tracks of a rank are g = rank + world*j: generate them one stride at a time
        for j, g in enumerate(ids) if world > 1 else ():
            eng.synth_tracks(buf.data_ptr() + j * samples * 4, int(g), 1, samples, args.seed)
        if world == 1:
            eng.synth_tracks(buf.data_ptr(), int(ids[0]), len(ids), samples, args.seed)
        ok = sh.add(buf.data_ptr(), off_full[:len(ids) + 1], [int(g) for g in ids], device=True)
        assert ok.all()
    eng.index_commit()
    barrier()
    t_build = time.perf_counter() - t0
    stats = eng.index_stats()
    log(f"[identify] rank {rank}: shard {stats} built in {t_build:.1f} s")
    del buf

    # ---- queries (identical on every rank: same seed)
    def make_queries(n_queries: int, salt: int):
        """n_queries noisy 5 s excerpts -> (windows tensor [Q, 3, 56000] on the device, true track, true start sample)"""
        rng = np.random.default_rng(args.seed + 10**6 + salt)
        q_track = rng.integers(0, args.tracks, n_queries)
        q_start = rng.integers(0, samples - 80000 + 1, n_queries)
        gen = torch.Generator(device=dev); gen.manual_seed(args.seed + 7 + salt)
        wins = torch.empty((n_queries, 3, 56000), dtype=torch.float32, device=dev)
        for c0 in range(0, n_queries, 2048):                        # bounded scratch: 2048 whole tracks at a time
            c1 = min(c0 + 2048, n_queries)
            torch.cuda.synchronize()          # the engine writes `src` on its own stream: torch must be done with the last one
            src = torch.empty((c1 - c0) * samples, dtype=torch.float32, device=dev)
            for j in range(c0, c1):
                eng.synth_tracks(src.data_ptr() + (j - c0) * samples * 4, int(q_track[j]), 1, samples, args.seed)
            eng.sync()
            idx = torch.from_numpy(q_start[c0:c1]).to(dev)[:, None] + torch.arange(80000, device=dev)[None, :]
            clips = src.view(c1 - c0, samples).gather(1, idx)
            del src, idx
            p_sig = clips.pow(2).mean(dim=1, keepdim=True)
            noise = torch.randn(clips.shape, generator=gen, device=dev) * torch.sqrt(p_sig / (10 ** (args.snr_db / 10)))
            clips = (clips + noise).clamp_(-1.0, 1.0)
            del noise
            for w, (a_, b_) in enumerate(WINDOWS):
                wins[c0:c1, w] = clips[:, a_:b_]
            del clips
        torch.cuda.synchronize()
        return wins, q_track, q_start

    def measure(n_queries: int, salt: int):
        wins, q_track, q_start = make_queries(n_queries, salt)
        n_win = n_queries * 3
        off = np.arange(n_win + 1, dtype=np.int64) * 56000

        def step():
            return sh.query(wins.data_ptr(), off, device=True)

        for _ in range(args.warmup):
            merged, n = step()
        barrier()
        eng.stage_times(); eng.set_stage_timing(True)
        launches0 = eng.launches
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()                                   # the step's own stream waits for / is waited on by this one
        for _ in range(args.steps):
            merged, n = step()
        ev1.record()
        torch.cuda.synchronize()
        dt_wall = time.perf_counter() - t0
        dt_dev = ev0.elapsed_time(ev1) * 1e-3
        if sh._xchg is not None:
            sh._xchg.check()
        stage = eng.stage_times(); eng.set_stage_timing(False)
        launches = eng.launches - launches0
        t_all = torch.tensor([dt_dev, dt_wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
        dt = float(t_all[0].item()) / args.steps        # CUDA events on the launching stream, max over ranks
        dt_wall = float(t_all[1].item()) / args.steps

        # accuracy in the reference's terms: sum aligned hashes per track over the three windows, top-1
        m = merged.cpu().numpy() if hasattr(merged, "cpu") else merged
        top1 = 0
        offs_ok = 0
        for q in range(n_queries):
            votes = {}
            first_off = {}
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
        if args.dump and rank == 0:                      # per-window digests and the first rows, to compare runs offline
            np.savez_compressed(f"{args.dump}_{n_queries}.npz", window_digest=np.bitwise_xor.reduce(row_sums & 0xFFFFFFFF, axis=1),
                                n_rows=(m[:, :, 0] >= 0).sum(axis=1), first_rows=m[:, :4, :].astype(np.int32))
        per_step = {k: v[0] / args.steps for k, v in stage.items()}
        del wins
        return {"queries": n_queries, "windows_per_step": n_win, "queries_per_s": n_queries / dt, "ms_per_step": dt * 1e3,
                "ms_per_step_wall": dt_wall * 1e3, "top1_accuracy": top1 / n_queries,
                "top1_offset_within_1_frame": offs_ok / max(top1, 1), "rows_digest": digest,
                "stage_ms_per_step": {k: round(v, 3) for k, v in per_step.items()}, "gpu_launches": int(launches)}

    sizes = [args.queries] + [int(x) for x in args.also_queries.split(",") if x.strip()]
    if args.exchange == "peer":
        sh.enable_peer_exchange(3 * max(sizes))
    results = [measure(nq, salt) for salt, nq in enumerate(sizes)]
    main_r = results[0]

    if rank == 0:
        print(json.dumps({
            "metric": f"queries/sec vs {args.tracks}-track index", "value": main_r["queries_per_s"], "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_r["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": f"identify {args.queries} x 5 s queries (3 x 3.5 s windows, {args.snr_db:g} dB SNR) "
                                   f"against {args.tracks} x {args.seconds:g} s tracks", "index_sharding": f"track g on rank g % {world}",
                       "windows_per_step": main_r["windows_per_step"]},
            "timed_region": "device-resident query PCM -> fingerprint -> fingerprint exchange -> probe/vote/rank -> row exchange -> "
                            "merged rows on every rank (CUDA events, max over ranks; wall clock beside it)",
            "exchange": args.exchange, "ms_per_step_wall": main_r["ms_per_step_wall"],
            "top1_accuracy": main_r["top1_accuracy"], "top1_offset_within_1_frame": main_r["top1_offset_within_1_frame"],
            "rows_digest": main_r["rows_digest"],
            "index": {"tracks_per_rank": stats["tracks"], "postings_per_rank": stats["postings"],
                      "segments_per_rank": stats["segments"], "device_bytes_per_rank": stats["device_bytes"],
                      "build_seconds": t_build},
            "stage_ms_per_step": main_r["stage_ms_per_step"], "gpu_launches": main_r["gpu_launches"],
            "other_batch_sizes": results[1:],
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()
