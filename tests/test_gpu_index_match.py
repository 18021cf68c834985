"""GPU parity for stages a7-a9: index order, votes, ranked rows -- bit-exact
against oracle/aid_oracle.c given the same hashes (BASELINE.json north_star)."""
import numpy as np
import pytest

from audio_ident_b200 import synth
from audio_ident_b200.engine import ragged

pytestmark = pytest.mark.gpu


def rows_equal(rows, n, ref):
    got = rows[:n]
    assert n == len(ref), (n, len(ref), got[:5], ref[:5])
    for name in ("count", "track", "offset", "q_first", "q_last"):
        assert np.array_equal(got[name], ref[name]), name


@pytest.fixture(scope="module")
def corpus(oracle):
    tracks = [synth.make_track(k, 20.0) for k in range(24)]
    fps = [oracle.fingerprint(x) for x in tracks]
    return tracks, fps


def build_oracle_index(oracle, fps):
    H = np.concatenate([h for h, _ in fps])
    T = np.concatenate([t for _, t in fps])
    TR = np.concatenate([np.full(len(h), k, np.uint32) for k, (h, _) in enumerate(fps)])
    return oracle.Index(H, TR, T)


def test_match_rows_bit_exact_given_same_hashes(fresh_index, oracle, corpus):
    eng = fresh_index
    tracks, fps = corpus
    hoff = np.concatenate([[0], np.cumsum([len(h) for h, _ in fps])])
    ok = eng.index_add_hashes(np.concatenate([h for h, _ in fps]), np.concatenate([t for _, t in fps]), hoff,
                              [oracle.num_frames(len(x)) for x in tracks], [f"t{k}" for k in range(len(tracks))])
    assert ok.all()
    ix = build_oracle_index(oracle, fps)
    qh, qt, qoff = [], [], [0]
    refs = []
    for q in range(40):
        pcm, _ = synth.make_query(tracks[q % len(tracks)], q, 5.0, 20.0)
        for a, b in ((0.0, 3.5), (0.75, 4.25), (1.5, 5.0)):
            h, t = oracle.fingerprint(pcm[int(a * 16000):int(b * 16000)])
            qh.append(h); qt.append(t); qoff.append(qoff[-1] + len(h))
            refs.append(ix.match(h, t))
    rows, n = eng.query_hashes(np.concatenate(qh), np.concatenate(qt), qoff)
    hits = 0
    for i, ref in enumerate(refs):
        rows_equal(rows[i], n[i], ref)
        hits += len(ref) > 0 and ref["track"][0] == (i // 3) % len(tracks)
    assert hits >= 0.9 * len(refs)


def test_query_from_pcm_end_to_end_top1(fresh_index, corpus):
    eng = fresh_index
    tracks, _ = corpus
    pcm, off = ragged(tracks)
    assert eng.index_add(pcm, off, [f"t{k}" for k in range(len(tracks))]).all()
    st = eng.index_stats()
    assert st["tracks"] == len(tracks) and st["segments"] == 1 and st["postings"] > 1000
    wins = []
    for q in range(30):
        clip, _ = synth.make_query(tracks[q % len(tracks)], 1000 + q, 5.0, 20.0)
        wins += [clip[int(a * 16000):int(b * 16000)] for a, b in ((0.0, 3.5), (0.75, 4.25), (1.5, 5.0))]
    qp, qo = ragged(wins)
    rows, n = eng.query(qp, qo)
    top1 = 0
    for q in range(30):
        votes = {}
        for w in range(3):
            for r in rows[3 * q + w][:n[3 * q + w]]:
                votes[int(r["track"])] = votes.get(int(r["track"]), 0) + int(r["count"])
        top1 += bool(votes) and max(votes, key=votes.get) == q % len(tracks)
    assert top1 >= 29
    assert eng.track_name(3) == "t3"


def test_many_tracks_two_segments_and_fat_buckets(fresh_index, oracle):
    """20k tiny tracks (crosses the 16384-track segment boundary), few distinct hashes (fat buckets, multi-round
    vote partitioning, exact-table overflow restart), duplicates of one track (more than 50 tied rows)."""
    eng = fresh_index
    rng = np.random.default_rng(5)
    n_tracks = 20000
    per = 12
    base_h = rng.integers(0, 1 << 24, 64).astype(np.uint32)
    H = base_h[rng.integers(0, 64, n_tracks * per)]
    T = np.tile(np.arange(per, dtype=np.uint32) * 3, n_tracks)
    # tracks 100..179 and 17000..17019 are copies of one another
    dup = np.concatenate([np.arange(100, 180), np.arange(17000, 17020)])
    for k in dup:
        H[k * per:(k + 1) * per] = H[100 * per:101 * per]
    # make (hash, t) unique inside a track as real fingerprints are
    H = (H & ~np.uint32(0xF)) | (np.tile(np.arange(per, dtype=np.uint32), n_tracks) & 0xF)
    hoff = np.arange(n_tracks + 1, dtype=np.int64) * per
    ok = eng.index_add_hashes(H, T, hoff, np.full(n_tracks, 100, np.int64), [f"n{k}" for k in range(n_tracks)])
    assert ok.all()
    assert eng.index_stats()["segments"] == 2
    TR = np.repeat(np.arange(n_tracks, dtype=np.uint32), per)
    ix = oracle.Index(H, TR, T)
    qs = [(H[100 * per:101 * per], T[100 * per:101 * per] + 7),
          (H[5 * per:6 * per], T[5 * per:6 * per]),
          (H[19999 * per:], T[19999 * per:] + 1),
          (np.tile(H[:per * 40], 1), np.tile(T[:per], 40))]
    qoff = np.concatenate([[0], np.cumsum([len(h) for h, _ in qs])])
    rows, n = eng.query_hashes(np.concatenate([h for h, _ in qs]), np.concatenate([t for _, t in qs]), qoff)
    for i, (h, t) in enumerate(qs):
        rows_equal(rows[i], n[i], ix.match(h, t))
    assert n[0] == 50


def test_delete_replace_and_persistence(fresh_index, oracle, corpus, tmp_path):
    eng = fresh_index
    tracks, fps = corpus
    pcm, off = ragged(tracks[:8])
    names = [f"t{k}" for k in range(8)]
    assert eng.index_add(pcm, off, names).all()
    clip, _ = synth.make_query(tracks[2], 77, 5.0, 20.0)
    rows, n = eng.query(clip, [0, len(clip)])
    assert n[0] >= 1 and rows[0]["track"][0] == 2
    before = rows[0][:n[0]].copy()
    eng.index_save(str(tmp_path))
    assert eng.index_delete("t2") and not eng.index_delete("t2") and not eng.index_delete("nope")
    rows, n = eng.query(clip, [0, len(clip)])
    assert all(r["track"] != 2 for r in rows[0][:n[0]])
    assert eng.index_stats()["tracks"] == 7
    # re-adding a name replaces the old entry under a new track number
    assert eng.index_add(tracks[2], [0, len(tracks[2])], ["t2"]).all()
    rows, n = eng.query(clip, [0, len(clip)])
    assert n[0] >= 1 and rows[0]["track"][0] == 8 and eng.track_name(8) == "t2"
    # reload the snapshot taken before the delete
    eng.index_load(str(tmp_path))
    rows, n = eng.query(clip, [0, len(clip)])
    assert np.array_equal(rows[0][:n[0]], before)
    # an emptied directory is an empty index (the reference's `make rebuild-index` wipes it: Makefile:84-93)
    empty = tmp_path / "empty"; empty.mkdir()
    eng.index_load(str(empty))
    assert eng.index_stats()["tracks"] == 0
    rows, n = eng.query(clip, [0, len(clip)])
    assert n[0] == 0


def test_empty_and_degenerate_queries(fresh_index, corpus):
    eng = fresh_index
    tracks, _ = corpus
    rows, n = eng.query(tracks[0][:16000], [0, 16000])       # empty index
    assert n[0] == 0
    eng.index_add(tracks[0], [0, len(tracks[0])], ["a"])
    pcm, off = ragged([np.zeros(0, np.float32), np.zeros(800, np.float32), np.zeros(16000, np.float32), tracks[0][:48000]])
    rows, n = eng.query(pcm, off)
    assert list(n[:3]) == [0, 0, 0] and n[3] >= 1 and rows[3]["track"][0] == 0 and rows[3]["offset"][0] == 0


def _tiny_corpus(n_tracks, per_track, hash_space, seed):
    """synthetic fingerprints: per_track (hash, t) pairs per track, hashes from a small space so that buckets are deep"""
    rng = np.random.default_rng(seed)
    h = rng.integers(0, hash_space, n_tracks * per_track, dtype=np.int64).astype(np.uint32)
    t = np.sort(rng.integers(0, 3000, (n_tracks, per_track)), axis=1).reshape(-1).astype(np.uint32)
    off = np.arange(n_tracks + 1, dtype=np.int64) * per_track
    return h, t, off


def test_segment_groups_keep_rows_bit_exact(fresh_index, oracle, tmp_path):
    """Full segments share a hash directory eight at a time (index.h SegGroup). 40,000 tracks = two full segments
    (grouped) + a partial one (its own table). Rows must equal the oracle's global index with and without grouping, and
    deletes, later adds (a third segment fills up and joins the group) and save + load must keep working. Deep buckets
    (~190 postings per hash) push the matcher through several sketch rounds and table restarts."""
    eng = fresh_index
    eng.index_set_grouping(True)
    NT, PER, SPACE = 40000, 24, 5000
    h, t, off = _tiny_corpus(NT, PER, SPACE, 7)
    # copies of track 5 in the other full segment and in the partial one: equal counts, ordered by track
    for g in (20000, 39000, 39001):
        h[g * PER:(g + 1) * PER] = h[5 * PER:6 * PER]; t[g * PER:(g + 1) * PER] = t[5 * PER:6 * PER]
    for c0 in range(0, NT, 8192):
        c1 = min(c0 + 8192, NT)
        ok = eng.index_add_hashes(h[off[c0]:off[c1]], t[off[c0]:off[c1]], off[c0:c1 + 1] - off[c0], [4000] * (c1 - c0),
                                  [str(g) for g in range(c0, c1)])
        assert ok.all()
    track_of = np.repeat(np.arange(NT, dtype=np.uint32), PER)
    ix = oracle.Index(h, track_of, t)

    rng = np.random.default_rng(11)
    picks = [5, 16383, 16384, 32767, 32768, 39999] + [int(g) for g in rng.integers(0, NT, 26)]
    qh, qt, qoff = [], [], [0]
    for g in picks:
        keep = rng.random(PER) < 0.8
        shift = int(t[g * PER:(g + 1) * PER].min())
        hh = h[g * PER:(g + 1) * PER][keep]; tt = (t[g * PER:(g + 1) * PER][keep] - shift).astype(np.uint32)
        # plus decoys so that a window has several hundred hashes
        dh = rng.integers(0, SPACE, 300, dtype=np.int64).astype(np.uint32); dt = rng.integers(0, 400, 300).astype(np.uint32)
        hh = np.concatenate([hh, dh]); tt = np.concatenate([tt, dt])
        o = np.argsort(tt, kind="stable")
        qh.append(hh[o]); qt.append(tt[o]); qoff.append(qoff[-1] + len(hh))
    QH, QT = np.concatenate(qh), np.concatenate(qt)

    def check(index, tomb=None):
        rows, n = eng.query_hashes(QH, QT, qoff)
        for i in range(len(picks)):
            rows_equal(rows[i], n[i], index.match(qh[i], qt[i], tombstone=tomb))
        return rows, n

    grouped, ng = check(ix)
    st = eng.index_stats()
    assert st["segments"] == 3 and st["segments_grouped"] == 2
    assert {int(x) for x in grouped[0]["track"][:4]} == {5, 20000, 39000, 39001} and ng[0] >= 4
    eng.index_set_grouping(False)
    plain, npl = check(ix)
    assert eng.index_stats()["segments_grouped"] == 0
    assert np.array_equal(npl, ng) and all(np.array_equal(plain[i][:ng[i]], grouped[i][:ng[i]]) for i in range(len(picks)))
    eng.index_set_grouping(True)

    # delete inside a grouped segment and inside the partial one
    assert eng.index_delete("20000") and eng.index_delete("39001")
    tomb = np.zeros(NT, np.uint8); tomb[[20000, 39001]] = 1
    rows, n = check(ix, tomb)
    assert {int(x) for x in rows[0]["track"][:2]} == {5, 39000}

    # later adds: the partial segment fills up and joins the group (3 members), a new partial one opens
    h2, t2, off2 = _tiny_corpus(10000, PER, SPACE, 8)
    h2[:PER] = h[5 * PER:6 * PER]; t2[:PER] = t[5 * PER:6 * PER]
    assert eng.index_add_hashes(h2, t2, off2, [4000] * 10000, [f"x{g}" for g in range(10000)]).all()
    H3 = np.concatenate([h, h2]); T3 = np.concatenate([t, t2])
    TR3 = np.concatenate([track_of, np.repeat(np.arange(NT, NT + 10000, dtype=np.uint32), PER)])
    ix3 = oracle.Index(H3, TR3, T3)
    tomb3 = np.zeros(NT + 10000, np.uint8); tomb3[[20000, 39001]] = 1
    rows, n = check(ix3, tomb3)
    assert 40000 in rows[0]["track"][:3]
    st = eng.index_stats()
    assert st["segments"] == 4 and st["segments_grouped"] == 3

    # persistence keeps the entries; the groups are derived data and come back on load
    eng.index_save(str(tmp_path))
    eng.index_clear()
    eng.index_load(str(tmp_path))
    assert eng.index_stats()["segments_grouped"] == 3
    rows2, n2 = check(ix3, tomb3)
    assert np.array_equal(n2, n) and all(np.array_equal(rows2[i][:n[i]], rows[i][:n[i]]) for i in range(len(picks)))
