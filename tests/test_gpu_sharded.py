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
