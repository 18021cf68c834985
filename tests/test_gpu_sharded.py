"""Sharded identification with the CUDA engine as the backend: two engines on one GPU stand in for two ranks
(each holds tracks g % 2 == r); their row blocks go through the same merge the ranks run after the all-gather.
Merged rows must equal one unsharded engine and the oracle. (The exchange itself is covered on CPU with gloo in
tests/test_sharded_cpu.py and by the N>1 runs of bench_identify.py on multi-GPU boxes.)"""
import numpy as np
import pytest

from audio_ident_b200 import sharded, synth
from audio_ident_b200.engine import Engine, ragged

pytestmark = pytest.mark.gpu


def test_two_shards_equal_one_index(engine, oracle):
    tracks = [synth.make_track(700 + k, 10.0) for k in range(10)]
    tracks.append(tracks[3].copy()); tracks.append(tracks[3].copy())          # duplicates landing on both shards
    engine.index_clear()
    pcm, off = ragged(tracks)
    assert engine.index_add(pcm, off, [str(g) for g in range(len(tracks))]).all()
    wins = []
    for q in range(8):
        clip, _ = synth.make_query(tracks[q], 50 + q, 5.0, 20.0)
        wins += [clip[:56000], clip[12000:68000], clip[24000:]]
    qp, qo = ragged(wins)
    rows1, n1 = engine.query(qp, qo)
    single = sharded.rows_to_array(rows1, n1, np.arange(len(tracks)))
    blocks = []
    for r in range(2):
        with Engine(0) as e2:
            sh = sharded.ShardedIdentifier(e2, r, 2)
            mine = sh.my_tracks(len(tracks))
            p2, o2 = ragged([tracks[g] for g in mine])
            assert sh.add(p2, o2, mine).all()
            rows, n = e2.query(qp, qo)
            blocks.append(sharded.rows_to_array(rows, n, np.asarray(sh.to_global)))
    merged, n = sharded.merge_rows(np.stack(blocks))
    assert np.array_equal(n, n1)
    for q in range(len(wins)):
        assert np.array_equal(merged[q, :n[q]], single[q, :n[q]]), q
    hits = sum(int(n[3 * q] > 0 and merged[3 * q, 0, 1] in (q, 10, 11) ) for q in range(8))
    assert hits == 8
    engine.index_clear()


def test_device_resident_query_path_equals_host_path(engine, oracle):
    """ShardedIdentifier.query(device=True) (fingerprint_dev -> aid_match_dev -> GPU merge, no host round trip)
    against the plain host-buffer query on the same index."""
    torch = pytest.importorskip("torch")
    tracks = [synth.make_track(800 + k, 8.0) for k in range(6)]
    engine.index_clear()
    pcm, off = ragged(tracks)
    sh = sharded.ShardedIdentifier(engine, 0, 1, device=torch.device("cuda", 0))
    assert sh.add(pcm, off, list(range(len(tracks)))).all()
    wins = [tracks[k % 6][4000 * k:4000 * k + 56000] for k in range(9)] + [np.zeros(300, np.float32)]
    qp, qo = ragged(wins)
    rows, n = engine.query(qp, qo)
    ref = sharded.rows_to_array(rows, n, np.arange(len(tracks)))
    d = torch.from_numpy(qp).cuda()
    merged, nn = sh.query(d.data_ptr(), qo, device=True)
    torch.cuda.synchronize()
    assert np.array_equal(nn.cpu().numpy(), n)
    m = merged.cpu().numpy()
    for q in range(len(wins)):
        assert np.array_equal(m[q, :n[q]], ref[q, :n[q]]), q
    assert (n[:9] >= 1).all() and n[9] == 0
    engine.index_clear()


def _device_fingerprints(engine, torch, qp, qo):
    """query windows -> (hash, t_anchor, hash_off u32) as torch tensors that outlive the engine's buffers"""
    n = len(qo) - 1
    d = torch.from_numpy(qp).cuda()
    res = engine.fingerprint_dev(d.data_ptr(), qo)
    off = torch.zeros(n + 1, dtype=torch.int32, device="cuda")
    engine.copy_device(off, res.d_hash_off, 4 * (n + 1))
    engine.sync()
    total = int(off[-1].item())
    h = torch.zeros(max(total, 1), dtype=torch.int32, device="cuda")
    t = torch.zeros(max(total, 1), dtype=torch.int32, device="cuda")
    engine.copy_device(h, res.d_hash, 4 * total)
    engine.copy_device(t, res.d_t_anchor, 4 * total)
    engine.sync()
    return h, t, off


def test_fused_rank_exchange_equals_one_index(engine):
    """aid_match_exchange_dev: three engines on one GPU stand in for three ranks (tracks g % 3 == r); k_rank stores
    its rows into every rank's window, k_merge_blocks merges.
    Every rank must end up with the rows of one unsharded index, over several epochs (the windows are
    double-buffered by epoch parity)."""
    torch = pytest.importorskip("torch")
    tracks = [synth.make_track(900 + k, 10.0) for k in range(9)]
    tracks += [tracks[2].copy(), tracks[2].copy(), tracks[2].copy()]          # duplicates: one on each rank
    engine.index_clear()
    pcm, off = ragged(tracks)
    assert engine.index_add(pcm, off, [str(g) for g in range(len(tracks))]).all()
    W = 3
    ranks = []
    try:
        for r in range(W):
            e = Engine(0)
            sh = sharded.ShardedIdentifier(e, r, W, device=torch.device("cuda", 0))
            mine = sh.my_tracks(len(tracks))
            p2, o2 = ragged([tracks[g] for g in mine])
            assert sh.add(p2, o2, mine).all()
            e.index_commit()
            ranks.append((e, sh))
        for epoch in range(3):
            wins = []
            for q in range(6):
                clip, _ = synth.make_query(tracks[(q + epoch) % 9], 70 + q + 10 * epoch, 5.0, 20.0)
                wins += [clip[:56000], clip[12000:68000], clip[24000:]]
            wins.append(np.zeros(200, np.float32))                               # a window without fingerprints
            qp, qo = ragged(wins)
            n = len(wins)
            rows1, n1 = engine.query(qp, qo)
            single = sharded.rows_to_array(rows1, n1, np.arange(len(tracks)))
            h, t, hoff = _device_fingerprints(engine, torch, qp, qo)
            if epoch == 0:
                for e, sh in ranks:
                    sh.enable_peer_exchange(64, connect=False)
                sharded.connect_local([sh for _, sh in ranks])
            outs = []
            streams = [torch.cuda.Stream() for _ in ranks]
            for (e, sh), st in zip(ranks, streams):                                # nobody waits on the host in between
                rows = torch.empty((n, 50, 5), dtype=torch.int32, device="cuda")
                nrows = torch.empty(n, dtype=torch.int32, device="cuda")
                tmap = torch.tensor(sh.to_global, dtype=torch.int32, device="cuda")
                e.match_exchange_dev(sh._xchg, h, t, hoff, None, None, n, tmap, len(sh.to_global), rows, nrows, 50,
                                     st.cuda_stream)
                outs.append((rows, nrows, tmap))
            torch.cuda.synchronize()
            for (e, sh), (rows, nrows, _) in zip(ranks, outs):
                sh._xchg.check()
                nn = nrows.cpu().numpy()
                assert np.array_equal(nn, n1)
                m = rows.cpu().numpy().astype(np.int64)
                for q in range(n):
                    assert np.array_equal(m[q, :nn[q]], single[q, :nn[q]]), (epoch, sh.rank, q)
            assert (n1[:18] >= 1).all() and n1[18] == 0
            # the whole step as one engine call per rank (aid_identify_exchange_dev): each rank fingerprints a third of
            # the windows, fingerprints and rows travel through the windows; callers sit on separate streams so that
            # no rank's stream waits for another rank's on the host side
            d_pcm = torch.from_numpy(qp).cuda()
            fused = []
            for (e, sh), st in zip(ranks, streams):
                with torch.cuda.stream(st):
                    fused.append(sh.query(d_pcm.data_ptr(), qo, device=True, check=False))   # ranks share a process: nobody may wait
            torch.cuda.synchronize()
            for (e, sh), (merged, nn_) in zip(ranks, fused):
                sh._xchg.check()
                assert np.array_equal(nn_.cpu().numpy(), n1)
                m = merged.cpu().numpy()
                for q in range(n):
                    assert np.array_equal(m[q, :n1[q]], single[q, :n1[q]]), (epoch, sh.rank, q, "fused step")
                assert (m[np.arange(50)[None, :] >= n1[:, None]] == -1).all()
            if epoch == 0:                          # query 2 is the triplicated track: rows from all three ranks interleave
                assert n1[6] >= 4 and {int(x) for x in single[6, :4, 1]} == {2, 9, 10, 11}
    finally:
        for e, _ in ranks:
            e.close()
        engine.index_clear()


def test_fused_exchange_single_rank_and_timeout(engine):
    """world 1: the fused path equals aid_match_dev. world 2 with a silent peer: the merge gives up after the
    timeout, reports AID_E_TIMEOUT and marks the windows (-1) instead of inventing rows."""
    torch = pytest.importorskip("torch")
    from audio_ident_b200.engine import EngineError
    tracks = [synth.make_track(950 + k, 8.0) for k in range(4)]
    engine.index_clear()
    pcm, off = ragged(tracks)
    assert engine.index_add(pcm, off, [str(g) for g in range(4)]).all()
    wins = [tracks[k % 4][3000 * k:3000 * k + 56000] for k in range(5)]
    qp, qo = ragged(wins)
    n = len(wins)
    h, t, hoff = _device_fingerprints(engine, torch, qp, qo)
    ref = torch.empty((n, 50, 5), dtype=torch.int32, device="cuda"); ref_n = torch.empty(n, dtype=torch.int32, device="cuda")
    engine.match_dev(h, t, hoff, None, None, n, ref, ref_n)
    x1 = engine.exchange(0, 1, 16)
    rows = torch.empty((n, 50, 5), dtype=torch.int32, device="cuda"); nrows = torch.empty(n, dtype=torch.int32, device="cuda")
    engine.match_exchange_dev(x1, h, t, hoff, None, None, n, None, 0, rows, nrows)
    engine.sync(); torch.cuda.synchronize()
    x1.check()
    nn = nrows.cpu().numpy()
    assert np.array_equal(nn, ref_n.cpu().numpy()) and (nn >= 1).all()
    a, b = rows.cpu().numpy(), ref.cpu().numpy()
    for q in range(n):
        assert np.array_equal(a[q, :nn[q]], b[q, :nn[q]])
    x1.close()

    # world 2 where rank 1's shard is empty: it still publishes its (empty) blocks and both ranks see rank 0's rows
    with Engine(0) as empty:
        xa, xb = engine.exchange(0, 2, 16), empty.exchange(1, 2, 16)
        xa.connect_local([xa, xb]); xb.connect_local([xa, xb])
        rows_b = torch.empty((n, 50, 5), dtype=torch.int32, device="cuda"); nrows_b = torch.empty(n, dtype=torch.int32, device="cuda")
        sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
        engine.match_exchange_dev(xa, h, t, hoff, None, None, n, None, 0, rows, nrows, 50, sa.cuda_stream)
        empty.match_exchange_dev(xb, h, t, hoff, None, None, n, None, 0, rows_b, nrows_b, 50, sb.cuda_stream)
        torch.cuda.synchronize()
        xa.check(); xb.check()
        for rr, nr_ in ((rows, nrows), (rows_b, nrows_b)):
            assert np.array_equal(nr_.cpu().numpy(), nn)
            a = rr.cpu().numpy()
            for q in range(n):
                assert np.array_equal(a[q, :nn[q]], b[q, :nn[q]])
        xa.close(); xb.close()

    # the whole-step call on one rank equals the staged calls; a window too small for the fingerprints is reported
    d_pcm = torch.from_numpy(qp).cuda()
    xs = engine.exchange(0, 1, 16)
    engine.identify_exchange_dev(xs, d_pcm.data_ptr(), qo, None, 0, rows, nrows)
    engine.sync(); torch.cuda.synchronize()
    xs.check()
    assert np.array_equal(nrows.cpu().numpy(), nn)
    a = rows.cpu().numpy()
    for q in range(n):
        assert np.array_equal(a[q, :nn[q]], b[q, :nn[q]])
    xs.close()
    tiny = engine.exchange(0, 1, 16, max_hashes_per_rank=8)
    engine.identify_exchange_dev(tiny, d_pcm.data_ptr(), qo, None, 0, rows, nrows)
    engine.sync()
    with pytest.raises(EngineError) as ei:
        tiny.check()
    assert ei.value.status == -3 and (nrows.cpu().numpy() == 0).all()
    tiny.close()

    x2 = engine.exchange(0, 2, 16)
    peer = engine.exchange(1, 2, 16)                   # exists, never publishes
    x2.connect_local([x2, peer]); peer.connect_local([x2, peer])
    x2.set_timeout_ms(100)
    engine.match_exchange_dev(x2, h, t, hoff, None, None, n, None, 0, rows, nrows)
    engine.sync()
    with pytest.raises(EngineError) as ei:
        x2.check()
    assert ei.value.status == -9
    assert (nrows.cpu().numpy() == -1).all()
    x2.close(); peer.close()
    engine.index_clear()


def test_host_entry_pipelines_its_uploads_and_keeps_the_rows(engine):
    """aid_identify_exchange_host / _windows_host cut a batch of >= 128 windows into up to eight parts whose uploads overlap the
    previous part's step (csrc/exchange.cu identify_host): rows and counts must equal the device-resident single step,
    for separate windows and for overlapping windows of shared clips, also when only a slice of the rows is returned."""
    torch = pytest.importorskip("torch")
    tracks = [synth.make_track(1200 + k, 12.0) for k in range(12)]
    engine.index_clear()
    pcm, off = ragged(tracks)
    sh = sharded.ShardedIdentifier(engine, 0, 1, device=torch.device("cuda", 0))
    assert sh.add(pcm, off, list(range(len(tracks)))).all()
    engine.index_commit()
    sh.enable_peer_exchange(512, connect=True)
    try:
        clips = [synth.make_query(tracks[q % 12], 500 + q, 5.0, 20.0)[0] for q in range(101)]
        wins = []
        for c in clips:
            wins += [c[:56000], c[12000:68000], c[24000:]]
        wins[7] = np.zeros(300, np.float32)                                   # no fingerprints
        qp, qo = ragged(wins)
        n = len(wins)
        assert n >= 256
        d_pcm = torch.from_numpy(qp).cuda()
        ref, ref_n = sh.query(d_pcm.data_ptr(), qo, device=True)
        ref, ref_n = ref.cpu().numpy(), ref_n.cpu().numpy()
        rows, nr = sh.query_host(qp, qo)
        assert np.array_equal(nr, ref_n)
        for q in range(n):
            got = np.stack([rows[q, :nr[q]][f] for f in rows.dtype.names], axis=1).astype(np.int64) if nr[q] else np.zeros((0, 5), np.int64)
            assert np.array_equal(got, ref[q, :nr[q]]), q
        rows2, nr2 = sh.query_host(qp, qo, rows_first=100, rows_count=57)
        assert np.array_equal(nr2, ref_n[100:157])
        # the same windows as offsets into ONE upload of every clip
        cp, co = ragged(clips)
        begin = np.concatenate([[co[i], co[i] + 12000, co[i] + 24000] for i in range(len(clips))]).astype(np.int64)
        end = np.concatenate([[co[i] + 56000, co[i] + 68000, co[i + 1]] for i in range(len(clips))]).astype(np.int64)
        rows3, nr3 = sh.query_host(cp, (begin, end))
        keep = np.arange(n) != 7                                              # window 7 was replaced by silence above
        assert np.array_equal(nr3[keep], ref_n[keep])
        for q in np.flatnonzero(keep):
            got = np.stack([rows3[q, :nr3[q]][f] for f in rows3.dtype.names], axis=1).astype(np.int64)
            assert np.array_equal(got, ref[q, :nr3[q]]), q
    finally:
        if sh._xchg is not None:
            sh._xchg.close()
        engine.index_clear()


def test_host_entry_error_in_a_later_part_leaves_the_engine_usable(engine):
    """A window longer than a vote window may be (AID_E_TOO_LONG) that sits in a later part of a pipelined batch: the call
    fails as a whole, nothing hangs, and the next call on the same exchange works."""
    torch = pytest.importorskip("torch")
    from audio_ident_b200.engine import EngineError
    tracks = [synth.make_track(1300 + k, 8.0) for k in range(4)]
    engine.index_clear()
    pcm, off = ragged(tracks)
    sh = sharded.ShardedIdentifier(engine, 0, 1, device=torch.device("cuda", 0))
    assert sh.add(pcm, off, list(range(len(tracks)))).all()
    engine.index_commit()
    sh.enable_peer_exchange(512, connect=True)
    try:
        wins = [tracks[k % 4][1000 * (k % 7):1000 * (k % 7) + 56000] for k in range(200)]
        good, good_off = ragged(wins)
        rows_ok, n_ok = sh.query_host(good, good_off)
        assert (n_ok >= 1).all()
        bad = list(wins)
        bad[150] = np.zeros(270 * 16000, np.float32)                          # 33,742 frames > AID_QUERY_MAX_FRAMES
        bp, bo = ragged(bad)
        with pytest.raises(EngineError) as ei:
            sh.query_host(bp, bo)
        assert ei.value.status == -4                                          # AID_E_TOO_LONG
        rows2, n2 = sh.query_host(good, good_off)
        assert np.array_equal(n2, n_ok)
    finally:
        if sh._xchg is not None:
            sh._xchg.close()
        engine.index_clear()
