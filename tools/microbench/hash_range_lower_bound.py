#!/usr/bin/env python
"""Lower bound on what HASH-RANGE sharding of the identification index (BASELINE.json north_star's wording) would cost
per step, measured with library stand-ins -- so that the choice of TRACK sharding rests on a
measurement and not on an argument.

With the index split by hash range every rank sees, for every query window, only the votes of its hash range: a
(track, offset) histogram is spread over all ranks, no rank can apply the >= 6 threshold, and every partial vote has to
travel to the rank that owns its track before anything can be counted. Per step that is, at the very least:
  1. an all-to-all of the partial votes as (window, track, offset) keys (8 B each), and
  2. a reduce-by-key of what arrives (a sort, here torch.sort as the stand-in for the best case),
on top of the same probing work the track-sharded step does. The vote count is not assumed: the matcher counts the
postings it touches (bench.py: identify.indexes[].runs[].k_match.postings_touched_per_step; 628,838,252 for 12,288
windows against 1,000,000 tracks).

  torchrun --nproc-per-node N tools/microbench/hash_range_lower_bound.py --votes 628838252 --step-ms <track-sharded ms at N>
"""
import argparse
import json
import os

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--votes", type=int, default=628838252, help="postings touched per step over ALL ranks (measured)")
    ap.add_argument("--step-ms", type=float, default=0.0, help="the whole track-sharded step at this N, for comparison")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    per_rank = args.votes // world                      # votes a rank produces from its hash range = votes it receives
    per_pair = per_rank // world
    gen = torch.Generator(device=dev); gen.manual_seed(rank)
    send = torch.randint(0, 1 << 62, (per_pair * world,), dtype=torch.int64, device=dev, generator=gen)
    recv = torch.empty_like(send)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t_x, t_s = [], []
    for _ in range(args.reps + 2):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ev[0].record()
        if world > 1:
            dist.all_to_all_single(recv, send)
        else:
            recv.copy_(send)
        ev[1].record()
        keys, _ = torch.sort(recv)                      # the cheapest possible reduce-by-key: one sort of what arrived
        ev[2].record()
        torch.cuda.synchronize()
        t_x.append(ev[0].elapsed_time(ev[1])); t_s.append(ev[1].elapsed_time(ev[2]))
        del keys
    t = torch.tensor([min(t_x[2:]), min(t_s[2:])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        x, s_ = float(t[0]), float(t[1])
        print(json.dumps({"what": "lower bound on the extra per-step cost of hash-range sharding (library stand-ins)",
                          "n_gpus": world, "votes_per_step": args.votes, "votes_per_rank": per_rank,
                          "all_to_all_ms": x, "all_to_all_gbs_per_rank": per_rank * 8 / x / 1e6 if x > 0 else None,
                          "sort_ms": s_, "extra_ms_per_step": x + s_, "track_sharded_step_ms": args.step_ms or None,
                          "extra_over_whole_track_sharded_step": (x + s_) / args.step_ms if args.step_ms else None}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
