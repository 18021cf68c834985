"""Size-independent properties at (a slice of) BASELINE.json's full sizes, where the CPU oracle is too slow to be the
checker: device-generated 30 s tracks by the thousand (configs[1]), self-identification against the index built from
them (configs[2]). Every property follows from the specification in include/aid_params.h."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N_TRACKS = 3000
SAMPLES = 30 * 16000


@pytest.fixture(scope="module")
def corpus(engine):
    d = engine.device_alloc(N_TRACKS * SAMPLES * 4)
    engine.synth_tracks(d, 0, N_TRACKS, SAMPLES, 42)
    engine.sync()
    yield d
    engine.device_free(d)


def fingerprint_all(engine, d, chunk):
    hs, ts, offs, base = [], [], [0], 0
    for k0 in range(0, N_TRACKS, chunk):
        n = min(chunk, N_TRACKS - k0)
        res = engine.fingerprint_dev(d + k0 * SAMPLES * 4, np.arange(n + 1, dtype=np.int64) * SAMPLES)
        off = engine.to_host(res.d_hash_off, n + 1, np.uint32).astype(np.int64)
        st = engine.to_host(res.d_status, n, np.int32)
        assert (st == 0).all()
        hs.append(engine.to_host(res.d_hash, int(off[-1]), np.uint32)); ts.append(engine.to_host(res.d_t_anchor, int(off[-1]), np.uint32))
        offs += list(base + off[1:]); base += int(off[-1])
    return np.concatenate(hs), np.concatenate(ts), np.asarray(offs, np.int64)


def test_deterministic_and_independent_of_batching(engine, corpus):
    a = fingerprint_all(engine, corpus, 1000)
    b = fingerprint_all(engine, corpus, 1000)
    c = fingerprint_all(engine, corpus, 333)          # different launch grouping, run lengths and unit boundaries
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    for x, y in zip(a, c):
        assert np.array_equal(x, y)


def test_kernel_generations_agree_on_3000_tracks(engine, corpus):
    """The round-1 kernels (scalar STFT, row-streaming peaks), the packed STFT with row-streaming peaks and the product
    configuration (packed STFT + group maxima for the peak kernel) must emit the same hashes, anchors and offsets for
    every one of 3,000 tracks (csrc/stft.cu, csrc/peaks.cu; aid_engine_set_kernels)."""
    try:
        engine.set_kernels(0, False)
        ref = fingerprint_all(engine, corpus, 1000)
        for variant, summary in ((5, False), (5, True), (7, True)):
            engine.set_kernels(variant, summary)
            got = fingerprint_all(engine, corpus, 750)
            for x, y in zip(ref, got):
                assert np.array_equal(x, y), (variant, summary)
    finally:
        engine.set_kernels(5, True)


def test_every_hash_obeys_the_specification(engine, corpus):
    h, t, off = fingerprint_all(engine, corpus, 1500)
    f1, f2, dt = (h >> 15).astype(np.int64), ((h >> 6) & 511).astype(np.int64), (h & 63).astype(np.int64)
    assert (h >> 24 == 0).all()
    assert ((dt >= 2) & (dt <= 33)).all()
    df = np.abs(f1 - f2)
    assert ((df >= 1) & (df <= 128)).all()
    assert (f1 >= 9).all() and (f2 >= 9).all()
    assert (t.astype(np.int64) + dt < 3743).all()                          # the target frame exists
    per = np.diff(off)
    assert per.min() > 300 and per.max() < 3000                            # ~800 hashes per 30 s track
    track = np.repeat(np.arange(N_TRACKS), per)
    # anchors in (t, f) order inside a track; at most AID_FANOUT = 8 hashes per anchor
    akey = (track.astype(np.int64) << 40) | (t.astype(np.int64) << 9) | f1
    assert (np.diff(akey) >= 0).all()
    _, counts = np.unique(akey, return_counts=True)
    assert counts.max() <= 8
    # (hash, t_anchor) is unique inside a track (what makes vote counts bounded by the query length)
    full = (track.astype(np.int64) << 42) | (t.astype(np.int64) << 24) | h
    assert len(np.unique(full)) == len(full)


def test_every_track_identifies_itself(engine, corpus):
    engine.index_clear()
    names = [f"s{k}" for k in range(N_TRACKS)]
    for k0 in range(0, N_TRACKS, 1000):
        n = min(1000, N_TRACKS - k0)
        assert engine.index_add(corpus + k0 * SAMPLES * 4, np.arange(n + 1, dtype=np.int64) * SAMPLES, names[k0:k0 + n], device=True).all()
    st = engine.index_stats()
    assert st["tracks"] == N_TRACKS and st["segments"] == 1
    # query: 8 s from the middle of every track, straight from the device buffer
    q_len, q_start = 8 * 16000, 10 * 16000
    h, t, off = fingerprint_all(engine, corpus, 1500)
    qh, qt, qo = [], [], [0]
    for k in range(N_TRACKS):
        a, b = off[k], off[k + 1]
        sel = (t[a:b] >= q_start // 128) & (t[a:b] < (q_start + q_len) // 128 - 45)
        qh.append(h[a:b][sel]); qt.append(t[a:b][sel] - q_start // 128); qo.append(qo[-1] + int(sel.sum()))
    rows, n = engine.query_hashes(np.concatenate(qh), np.concatenate(qt), qo)
    assert (n >= 1).all()
    assert np.array_equal(rows["track"][:, 0], np.arange(N_TRACKS))
    assert (rows["offset"][:, 0] == q_start // 128).all()
    # the winning count is exactly the number of query hashes (every one of them aligns)
    assert np.array_equal(rows["count"][:, 0], np.diff(qo))
    engine.index_clear()
