"""The real multi-process exchange: two processes, one per GPU, each with its own engine and half of the tracks, swap the
64-byte handles of aid_exchange_handle through a pipe, map each other's receive window with aid_exchange_connect
(cudaIpcOpenMemHandle over NVLink) and run the fused identification step (aid_identify_exchange_dev and _host). Both must
end up with the rows of one unsharded index. Everything the torchrun path does, minus NCCL.

Needs TWO GPUs (skipped otherwise): the ranks' kernels wait for each other's flags on the device, and processes that
time-slice ONE GPU are not guaranteed to run at the same time (B200_PROFILING.md warns of Xid 109 for exactly that).
Run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_ipc_exchange.py -m gpu`; the log of that run is committed
under profiles/. (Ranks that share one process and one GPU -- tests/test_gpu_sharded.py -- use connect_local.)"""
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _corpus():
    from audio_ident_b200 import synth
    tracks = [synth.make_track(600 + k, 12.0) for k in range(10)]
    wins = []
    for q in range(8):
        clip, _ = synth.make_query(tracks[(3 * q) % 10], 40 + q, 5.0, 20.0)
        wins += [clip[:56000], clip[12000:68000], clip[24000:]]
    wins.append(np.zeros(300, np.float32))               # a window without fingerprints
    return tracks, wins


def _rank_main(rank, world, conn, out_q):
    sys.path.insert(0, ROOT)
    import torch
    from audio_ident_b200 import sharded
    from audio_ident_b200.engine import Engine, ragged
    try:
        tracks, wins = _corpus()
        torch.cuda.set_device(rank)
        eng = Engine(rank)                                  # one process per GPU
        sh = sharded.ShardedIdentifier(eng, rank, world, device=torch.device("cuda", rank))
        mine = sh.my_tracks(len(tracks))
        p, o = ragged([tracks[g] for g in mine])
        assert sh.add(p, o, mine).all()
        eng.index_commit()
        sh.enable_peer_exchange(64, connect=False)
        conn.send(sh._xchg.handle())                      # 64-byte cudaIpcMemHandle_t to the peer ...
        peer = conn.recv()                                # ... and the peer's back
        handles = [None, None]
        handles[rank], handles[1 - rank] = sh._xchg.handle(), peer
        sh._xchg.connect(handles)
        conn.send("connected"); assert conn.recv() == "connected"
        qp, qo = ragged(wins)
        d = torch.from_numpy(qp).cuda()
        res = []
        for _ in range(3):                                # three epochs: both parities of the double-buffered window
            merged, n = sh.query(d.data_ptr(), qo, device=True)          # check=True: raises on a failed exchange
            res.append((merged.cpu().numpy(), n.cpu().numpy()))
        n_win = len(wins)
        lo, hi = rank * n_win // world, (rank + 1) * n_win // world
        rows_h, n_h = sh.query_host(qp, qo, lo, hi - lo)  # host buffers in, this rank's slice of the rows out
        rows_h = np.stack([rows_h[name].astype(np.int64) for name in ("count", "track", "offset", "q_first", "q_last")], axis=2)
        conn.send("done"); assert conn.recv() == "done"   # nobody unmaps a window a peer may still store into
        out_q.put((rank, "ok", res, (lo, hi, rows_h.copy(), n_h.copy())))
        eng.close()
    except Exception as ex:                               # noqa: BLE001
        import traceback
        out_q.put((rank, "error", traceback.format_exc(), None))


def test_two_processes_two_gpus_cuda_ipc(engine):
    import ctypes
    from audio_ident_b200 import _lib, sharded
    if _lib.load().aid_device_count() < 2:
        pytest.skip("needs two GPUs: the ranks' kernels wait on each other and must not time-slice one device")
    from audio_ident_b200.engine import ragged
    tracks, wins = _corpus()
    engine.index_clear()
    p, o = ragged(tracks)
    assert engine.index_add(p, o, [str(g) for g in range(len(tracks))]).all()
    qp, qo = ragged(wins)
    rows1, n1 = engine.query(qp, qo)
    single = sharded.rows_to_array(rows1, n1, np.arange(len(tracks)))
    engine.index_clear()
    assert (n1[:24] >= 1).all() and n1[24] == 0

    ctx = mp.get_context("spawn")
    a, b = ctx.Pipe()
    out_q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, c, out_q)) for r, c in ((0, a), (1, b))]
    for pr in procs:
        pr.start()
    got = {}
    try:
        for _ in range(2):
            rank, status, res, host = out_q.get(timeout=240)
            assert status == "ok", res
            got[rank] = (res, host)
    finally:
        for pr in procs:
            pr.join(timeout=30)
            if pr.is_alive():
                pr.kill()
    for rank in (0, 1):
        res, (lo, hi, rows_h, n_h) = got[rank]
        for merged, n in res:
            assert np.array_equal(n, n1), rank
            for q in range(len(wins)):
                assert np.array_equal(merged[q, :n1[q]], single[q, :n1[q]]), (rank, q)
                assert (merged[q, n1[q]:] == -1).all()
        assert np.array_equal(n_h, n1[lo:hi])
        for q in range(lo, hi):
            assert np.array_equal(rows_h[q - lo, :n1[q]], single[q, :n1[q]]), (rank, q, "host path")
