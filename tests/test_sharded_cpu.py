"""N>1 identification path on CPU: world_size-2 (and 3) gloo processes, each holding the tracks g % P == rank in an
oracle-backed stand-in for the engine, exchange their row blocks with all_gather and merge. The merged rows must be
bit-identical to one unsharded oracle index over all tracks. (The CUDA engine plays the backend's part in
tests/test_gpu_sharded.py.)"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from audio_ident_b200 import sharded  # noqa: E402
from audio_ident_b200._lib import MATCH_ROW_DTYPE  # noqa: E402


class OracleBackend:
    """Engine-shaped object backed by the CPU oracle: just enough for ShardedIdentifier."""

    def __init__(self):
        self.h, self.t, self.tr, self.n = [], [], [], 0

    def index_add_hashes(self, h, t, hash_off, n_frames, names):
        for i in range(len(names)):
            a, b = int(hash_off[i]), int(hash_off[i + 1])
            self.h.append(np.asarray(h[a:b], np.uint32)); self.t.append(np.asarray(t[a:b], np.uint32))
            self.tr.append(np.full(b - a, self.n, np.uint32)); self.n += 1
        return np.ones(len(names), bool)

    def fingerprint(self, pcm, sample_off):
        from oracle import oracle
        sample_off = np.asarray(sample_off, np.int64)
        h, t, off, nh, npk, used = oracle.fingerprint_batch(np.asarray(pcm, np.float32)[sample_off[0]:], sample_off - sample_off[0], 1)
        return h, t, off, np.zeros(len(sample_off) - 1, np.int32)

    def query(self, pcm, sample_off, device=False):
        h, t, off, _ = self.fingerprint(pcm, sample_off)
        return self.query_hashes(h, t, off)

    def query_hashes(self, h, t, hash_off):
        from oracle import oracle
        ix = oracle.Index(np.concatenate(self.h), np.concatenate(self.tr), np.concatenate(self.t))
        nq = len(hash_off) - 1
        rows = np.zeros((nq, 50), MATCH_ROW_DTYPE); n = np.zeros(nq, np.int32)
        for q in range(nq):
            r = ix.match(h[hash_off[q]:hash_off[q + 1]], t[hash_off[q]:hash_off[q + 1]])
            rows[q, :len(r)] = r; n[q] = len(r)
        return rows, n


def make_corpus():
    """Tiny synthetic fingerprints with deliberate collisions: copies of tracks land on different ranks, so the
    merge has to break count ties by global track number, and more than 50 rows compete for the top-50."""
    rng = np.random.default_rng(3)
    n_tracks, per = 150, 14
    H = rng.integers(0, 1 << 24, (n_tracks, per)).astype(np.uint32)
    T = np.tile(np.arange(per, dtype=np.uint32) * 4, (n_tracks, 1))
    for g in range(20, 90):
        H[g] = H[20]                       # 70 copies of track 20 spread over every rank
    H[100, :8] = H[7, :8]                  # a partial copy: 8 aligned hashes
    queries = [(H[20], T[20] + 3), (H[7], T[7]), (H[120], T[120] + 9), (np.concatenate([H[5], H[6]]), np.concatenate([T[5], T[6] + 50]))]
    return H, T, queries


def worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    H, T, queries = make_corpus()
    sh = sharded.ShardedIdentifier(OracleBackend(), rank, world)
    mine = sh.my_tracks(len(H))
    per = H.shape[1]
    ok = sh.add_hashes(H[mine].reshape(-1), T[mine].reshape(-1), np.arange(len(mine) + 1) * per, np.full(len(mine), 100), mine)
    assert ok.all() and sh.to_global == mine
    qoff = np.concatenate([[0], np.cumsum([len(h) for h, _ in queries])])
    merged, n = sh.query_hashes(np.concatenate([h for h, _ in queries]), np.concatenate([t for _, t in queries]), qoff)
    q.put((rank, merged, n))
    dist.barrier()
    dist.destroy_process_group()


def pcm_worker(rank, world, port, q):
    """queries given as PCM: rank r fingerprints its slice of the windows, hashes are all-gathered (split path)"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from audio_ident_b200 import synth
    from oracle import oracle
    tracks = [synth.make_track(900 + k, 6.0) for k in range(6)]
    sh = sharded.ShardedIdentifier(OracleBackend(), rank, world)
    for g in sh.my_tracks(len(tracks)):
        h, t = oracle.fingerprint(tracks[g])
        sh.add_hashes(h, t, [0, len(h)], [oracle.num_frames(len(tracks[g]))], [g])
    wins = [tracks[k % 6][8000 * (k % 3):8000 * (k % 3) + 56000] for k in range(7)]      # 7 windows: uneven split
    pcm = np.concatenate(wins); off = np.concatenate([[0], np.cumsum([len(w) for w in wins])])
    a = sh.query(pcm, off, split_fingerprint=True)
    b = sh.query(pcm, off, split_fingerprint=False)
    q.put((rank, a, b))
    dist.barrier()
    dist.destroy_process_group()


def test_split_fingerprinting_equals_replicated(oracle):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=pcm_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, a, b in results:
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        assert (a[1] >= 1).all() and [int(a[0][k, 0, 1]) for k in range(7)] == [k % 6 for k in range(7)]
    assert np.array_equal(results[0][1][0], results[1][1][0])


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_rows_equal_unsharded_oracle(world, oracle):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    H, T, queries = make_corpus()
    n_tracks, per = H.shape
    ix = oracle.Index(H.reshape(-1), np.repeat(np.arange(n_tracks, dtype=np.uint32), per), T.reshape(-1))
    for rank, merged, n in results:
        for qi, (h, t) in enumerate(queries):
            ref = ix.match(h, t)
            assert n[qi] == len(ref), (rank, qi)
            for j, name in enumerate(("count", "track", "offset", "q_first", "q_last")):
                assert np.array_equal(merged[qi, :n[qi], j], ref[name].astype(np.int64)), (rank, qi, name)
    assert results[0][2][0] == 50              # 70 tied copies compete for 50 rows
    for r in results[1:]:
        assert np.array_equal(r[1], results[0][1]) and np.array_equal(r[2], results[0][2])


def test_single_rank_is_a_passthrough(oracle):
    H, T, queries = make_corpus()
    sh = sharded.ShardedIdentifier(OracleBackend(), 0, 1)
    per = H.shape[1]
    sh.add_hashes(H.reshape(-1), T.reshape(-1), np.arange(len(H) + 1) * per, np.full(len(H), 100), list(range(len(H))))
    h, t = queries[1]
    merged, n = sh.query_hashes(h, t, [0, len(h)])
    ix = oracle.Index(H.reshape(-1), np.repeat(np.arange(len(H), dtype=np.uint32), per), T.reshape(-1))
    ref = ix.match(h, t)
    assert n[0] == len(ref) and np.array_equal(merged[0, :n[0], 0], ref["count"])
