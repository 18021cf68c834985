"""At-scale parity, asserted ("close the parity gaps with tests, not bench prose"):

 (a) 512 device-generated 30 s tracks of the bench corpus, GPU vs the CPU oracle end to end: every track whose hashes
     differ is taken apart stage by stage and every one-sided peak must be a float near-tie (oracle/parity.py); the
     number of such tracks is bounded.
 (b) an index of 40,000 real fingerprints -- three segments, two of them sealed and sharing one group directory --
     probed with 312 noisy 3.5 s windows: rows (count, track, offset, q_first, q_last) bit-equal to oracle.Index.match
     over all 32 M postings.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SAMPLES = 480000


def _device_tracks(engine, first, n):
    d = engine.device_alloc(n * SAMPLES * 4)
    try:
        engine.synth_tracks(d, first, n, SAMPLES, 42)
        engine.sync()
        return engine.to_host(d, n * SAMPLES, np.float32)
    finally:
        engine.device_free(d)


def test_512_bench_tracks_against_the_oracle(engine, oracle):
    from oracle import parity
    n = 512
    pcm = _device_tracks(engine, 7000, n)
    off = np.arange(n + 1, dtype=np.int64) * SAMPLES
    gh, gt, goff, st = engine.fingerprint(pcm, off)
    assert (st == 0).all()
    rh, rt, roff, rnh, rnp, _ = oracle.fingerprint_batch(pcm, off, 0)
    differing = [i for i in range(n)
                 if not (np.array_equal(gh[goff[i]:goff[i + 1]], rh[roff[i]:roff[i + 1]]) and
                         np.array_equal(gt[goff[i]:goff[i + 1]], rt[roff[i]:roff[i + 1]]))]
    one_sided = 0
    for i in differing:
        k = parity.explain_track(engine, oracle, pcm[i * SAMPLES:(i + 1) * SAMPLES])     # raises on anything but a near-tie
        assert k >= 1, f"track {i}: hashes differ although the peak sets agree"
        one_sided += k
    # round 1 measured 9 of 1024 tracks with a flipped near-tie peak; a regression of the spectrogram's rounding error
    # shows up here first
    assert len(differing) <= 12, (len(differing), one_sided)
    assert one_sided <= 3 * max(len(differing), 1)
    assert int(goff[-1]) > 350 * n


def test_40k_track_index_rows_equal_oracle_match(engine, oracle):
    n_tracks, chunk = 40000, 1000
    engine.index_clear()
    hs, ts, lens = [], [], []
    d = engine.device_alloc(chunk * SAMPLES * 4)
    off = np.arange(chunk + 1, dtype=np.int64) * SAMPLES
    try:
        for c0 in range(0, n_tracks, chunk):
            engine.synth_tracks(d, c0, chunk, SAMPLES, 42)
            res = engine.fingerprint_dev(d, off)
            hoff = engine.to_host(res.d_hash_off, chunk + 1, np.uint32).astype(np.int64)
            st = engine.to_host(res.d_status, chunk, np.int32)
            assert (st == 0).all()
            hs.append(engine.to_host(res.d_hash, int(hoff[-1]), np.uint32))
            ts.append(engine.to_host(res.d_t_anchor, int(hoff[-1]), np.uint32))
            lens.append(np.diff(hoff))
            assert engine.index_add(d, off, [str(g) for g in range(c0, c0 + chunk)], device=True).all()
        engine.index_commit()
        stats = engine.index_stats()
        assert stats["tracks"] == n_tracks and stats["segments"] == 3 and stats["segments_grouped"] == 2
        H, T, L = np.concatenate(hs), np.concatenate(ts), np.concatenate(lens)
        assert stats["postings"] == len(H)
        ix = oracle.Index(H, np.repeat(np.arange(n_tracks, dtype=np.uint32), L), T)

        # 104 queries x three 3.5 s windows at a random sample offset, 20 dB SNR (bench_identify's query model)
        rng = np.random.default_rng(5)
        q_track = rng.integers(0, n_tracks, 104)
        wins = []
        for q, g in enumerate(q_track):
            engine.synth_tracks(d, int(g), 1, SAMPLES, 42)
            engine.sync()
            x = engine.to_host(d, SAMPLES, np.float32)
            s0 = int(rng.integers(0, SAMPLES - 80000 + 1))
            clip = x[s0:s0 + 80000].astype(np.float64)
            clip += rng.standard_normal(80000) * np.sqrt(np.mean(clip ** 2) / 100.0)
            clip = np.clip(clip, -1, 1).astype(np.float32)
            wins += [clip[0:56000], clip[12000:68000], clip[24000:80000]]
        pcm = np.concatenate(wins)
        woff = np.arange(len(wins) + 1, dtype=np.int64) * 56000
        qh, qt, qoff, qst = engine.fingerprint(pcm, woff)
        rows, n = engine.query(pcm, woff)
        hit = 0
        for w in range(len(wins)):
            ref = ix.match(qh[qoff[w]:qoff[w + 1]], qt[qoff[w]:qoff[w + 1]])
            assert n[w] == len(ref), (w, n[w], len(ref))
            for name in ("count", "track", "offset", "q_first", "q_last"):
                assert np.array_equal(rows[w][name][:n[w]], ref[name]), (w, name)
            hit += n[w] > 0 and rows[w]["track"][0] == q_track[w // 3]
        assert hit >= 0.95 * len(wins)
    finally:
        engine.device_free(d)
        engine.index_clear()
