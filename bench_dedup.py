#!/usr/bin/env python
"""bench_dedup.py -- SURVEY.md section 8(f)-4: content-duplicate checks against a resident Chromaprint store.

  python bench_dedup.py [--rows 1000000 --words 234 --queries 16 --steps 5 --warmup 3]

Workload: `rows` stored fingerprints of `words` 32-bit sub-fingerprints each (a 30 s track at Chromaprint's ~7.8
items/s) with durations inside the query's +-10 % window, so every row is a candidate (the worst case of
audio-ident-service/app/audio/dedup.py:192-212); each step checks `queries` new tracks, one scan per query, through
audio_ident_b200.dedup.ContentStore (host buffers in, host results out -- that is the e2e number; `value` uses the
CUDA-event time of the scan kernels alone). One JSON line on stdout: value = fingerprint rows compared per second,
roofline = algorithmic bytes of one scan (4 B per compared word + 20 B of row table per row) over the kernel
time, cpu_baseline = oracle/dedup_oracle.c (a C port of the reference's Python loop, all host threads, a bounded
sample of the rows) and `python_loop` = the reference's pure-Python arithmetic restated inline on a small sample.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
import uuid

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from bench import ClockSampler, measured_peaks, log  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--words", type=int, default=234)
    ap.add_argument("--queries", type=int, default=16)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu-rows", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    from audio_ident_b200 import dedup
    rng = np.random.default_rng(42)
    n, L = args.rows, args.words
    words = rng.integers(0, 2**32, n * L, dtype=np.uint64).astype(np.uint32)
    off = np.arange(n + 1, dtype=np.int64) * L
    dur = rng.uniform(28.0, 32.0, n)
    store = dedup.ContentStore(0)
    step_rows = 1 << 18
    for a in range(0, n, step_rows):
        b = min(n, a + step_rows)
        store.add_words([uuid.UUID(int=i) for i in range(a, b)], words[a * L:b * L], off[a:b + 1] - off[a], dur[a:b])

    def make_queries(seed):
        r = np.random.default_rng(seed)
        src = r.integers(0, n, args.queries)
        q = np.stack([words[s * L:(s + 1) * L] for s in src]).copy()
        flips = r.integers(0, L, (args.queries, L // 25))
        for i in range(args.queries):
            q[i, flips[i]] ^= np.uint32(1) << r.integers(0, 32, flips.shape[1]).astype(np.uint32)
        return src, q.reshape(-1), np.arange(args.queries + 1, dtype=np.int64) * L

    lo, hi = np.full(args.queries, 30.0 * 0.9), np.full(args.queries, 30.0 * 1.1)
    for w in range(args.warmup):
        store.scan_words(*make_queries(100 + w)[1:], lo, hi)
    sampler = ClockSampler(0); sampler.start()
    kernel_ms, wall = 0.0, 0.0
    ok = True
    for s in range(args.steps):
        src, qw, qo = make_queries(200 + s)
        t0 = time.perf_counter()
        rows, sims = store.scan_words(qw, qo, lo, hi)
        wall += time.perf_counter() - t0
        kernel_ms += store.last_scan_ms
        ok &= bool((rows == src).all() and (sims > 0.9).all())
    clocks = sampler.stop()
    per_scan_ms = kernel_ms / (args.steps * args.queries)
    compared = float(n) * args.queries * args.steps
    value = compared / (kernel_ms * 1e-3)
    e2e = compared / wall
    bytes_per_scan = n * (L * 4 + 20)
    peak, peak_src = measured_peaks()
    achieved = bytes_per_scan / (per_scan_ms * 1e-3) / 1e9

    out = {"metric": "fingerprint rows compared/sec (content-duplicate scan)", "value": value, "unit": "rows/s", "n_gpus": 1,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": kernel_ms / args.steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
           "config": {"workload": f"{args.queries} duplicate checks per step against {n} resident fingerprints of {L} words, every row a candidate",
                      "l2": f"store of {bytes_per_scan / 1e6:.0f} MB is larger than L2; streamed once per query"},
           "e2e": {"value": e2e, "unit": "rows/s", "h2d_bytes_per_step": int(args.queries * (L * 4 + 24)),
                   "d2h_bytes_per_step": int(args.queries * 16), "checks_per_s": args.queries * args.steps / wall},
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": None, "kernel": "k_dedup_scan", "ms_per_launch": per_scan_ms, "peak_source": peak_src},
           "gpu_launches": int(store.launches), "clocks": clocks, "top1_is_planted_row": ok}

    if not args.no_cpu:
        from oracle import oracle
        threads = len(os.sched_getaffinity(0))
        m = args.cpu_rows or min(n, 200_000)
        src, qw, qo = make_queries(300)
        qsel = slice(0, 2)
        oracle.dedup_scan(words[:m * L], off[:m + 1], dur[:m], qw[:2 * L], qo[:3], lo[qsel], hi[qsel], threads)
        t0 = time.perf_counter()
        reps = 0
        while time.perf_counter() - t0 < 5.0:
            oracle.dedup_scan(words[:m * L], off[:m + 1], dur[:m], qw[:2 * L], qo[:3], lo[qsel], hi[qsel], threads)
            reps += 1
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": m * 2 * reps / dt, "unit": "rows/s", "cores": threads, "kind": "port",
                               "sample": f"2 queries x the first {m} rows, {reps} repetitions (oracle/dedup_oracle.c, OpenMP)"}
        # the reference's arithmetic as it runs it: a Python loop with bin().count("1") (dedup.py:156-167)
        k = 200
        a1 = [int(x) for x in qw[:L]]
        t0 = time.perf_counter()
        for r in range(k):
            a2 = [int(x) for x in words[r * L:(r + 1) * L]]
            mb = 0
            for i in range(L):
                mb += 32 - bin((a1[i] ^ a2[i]) & 0xFFFFFFFF).count("1")
            _ = (mb / (L * 32)) * 1.0
        out["python_loop"] = {"value": k / (time.perf_counter() - t0), "unit": "rows/s", "cores": 1,
                              "sample": f"{k} rows, pre-parsed integers (the reference also re-parses both strings per row)"}
    store.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
