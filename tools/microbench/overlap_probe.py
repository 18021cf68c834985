#!/usr/bin/env python
"""Does the vote (k_match: latency / LSU bound) overlap with fingerprinting (k_stft_packed: FP32-pipe bound, all of an SM's
registers at 3 CTAs) when the two run on different streams of one GPU? Two engines on one device: A holds the index and
votes on fingerprints taken earlier, B fingerprints the same window batch. Prints the times alone, back to back on one
stream, and concurrently (with and without stream priority for the vote).  usage: overlap_probe.py [tracks=500000] [queries=4096]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    import bench_identify as bi
    from audio_ident_b200 import sharded
    from audio_ident_b200.engine import Engine
    tracks = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
    nq = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    ea, eb = Engine(0), Engine(0)
    cx = bi.Ctx(ea, 0, 1, dev)
    sh = sharded.ShardedIdentifier(ea, 0, 1, device=dev)
    print(f"building {tracks} tracks ...", flush=True)
    bi.build_shard(cx, sh, tracks)
    wins, q_track, q_start, _ = bi.make_queries(cx, tracks, nq, 1)
    n = nq * 3
    off = np.arange(n + 1, dtype=np.int64) * bi.WIN
    res = ea.fingerprint_dev(wins.data_ptr(), off)
    ea.sync()
    hoff = torch.from_numpy(ea.to_host(res.d_hash_off, n + 1, np.uint32).astype(np.int32)).to(dev)
    tot = int(hoff[-1].item())
    h = torch.from_numpy(ea.to_host(res.d_hash, tot, np.uint32).astype(np.int64)).to(dev).to(torch.int32)
    t = torch.from_numpy(ea.to_host(res.d_t_anchor, tot, np.uint32).astype(np.int64)).to(dev).to(torch.int32)
    rows = torch.empty((n, 50, 5), dtype=torch.int32, device=dev)
    nrows = torch.empty(n, dtype=torch.int32, device=dev)

    def run(label, f_stream, m_stream, reps=5):
        torch.cuda.synchronize()
        out = []
        for _ in range(reps):
            t0 = time.perf_counter()
            if f_stream is not None:
                eb.fingerprint_dev(wins.data_ptr(), off, f_stream.cuda_stream)
            if m_stream is not None:
                ea.match_dev(h, t, hoff, None, None, n, rows, nrows, 50, m_stream.cuda_stream)
            torch.cuda.synchronize()
            out.append((time.perf_counter() - t0) * 1e3)
        print(f"{label:44s} {min(out):7.3f} ms (median {sorted(out)[len(out) // 2]:7.3f})", flush=True)
        return min(out)

    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    lo, hi = torch.cuda.Stream(priority=0), torch.cuda.Stream(priority=-1)
    for _ in range(2):
        run("warm", s1, s1, 1)
    tf = run("fingerprint alone", s1, None)
    tm = run("vote alone", None, s2)
    ts = run("both, one stream (back to back)", s1, s1)
    tc = run("both, two streams", s1, s2)
    tp = run("both, vote on a high-priority stream", lo, hi)
    tq = run("both, fingerprint on the high-priority stream", hi, lo)
    print(f"sum {tf + tm:.3f}  serial {ts:.3f}  concurrent {tc:.3f} / {tp:.3f} / {tq:.3f}  max(alone) {max(tf, tm):.3f}")
    ea.close(); eb.close()


if __name__ == "__main__":
    main()
