"""The CPU oracle (oracle/aid_oracle.c) against (1) committed golden vectors for BASELINE.json configs[0],
(2) an independent numpy/scipy restatement of the same spec, (3) properties of the domain. CPU only.

The reference has no golden vectors for these stages ("parity unpinned"); the reference's
own known-answer tests for the boundary are replayed in tests/test_boundary_contract.py."""
import os

import numpy as np
import pytest

from audio_ident_b200 import synth
from oracle import np_oracle as npo

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "clip10_fingerprint.npz"))


def test_params_agree_everywhere(oracle):
    p = oracle.params()
    assert (p["nfft"], p["hop"], p["nbins"], p["sample_rate"]) == (npo.NFFT, npo.HOP, npo.NBINS, 16000)
    assert (p["peak_half_f"], p["peak_half_t"], p["peak_min_bin"]) == (npo.HALF_F, npo.HALF_T, npo.MIN_BIN)
    assert (p["dt_min"], p["dt_max"], p["df_min"], p["df_max"], p["fanout"]) == (2, 33, 1, 128, 8)
    assert (p["min_votes"], p["max_rows"], p["query_max_frames"], p["seg_tracks"]) == (6, 50, 32768, 16384)
    # the Python host code's copy of the same constants
    from audio_ident_b200 import _lib
    out = np.zeros(16, np.int32)
    import ctypes as C
    _lib.load().aid_get_params(out.ctypes.data_as(C.POINTER(C.c_int32)))
    assert dict(zip(oracle.PARAM_NAMES, (int(v) for v in out))) == p


def test_synthetic_clip_is_reproducible(clip10):
    assert len(clip10) == int(GOLD["n_samples"]) == 160000
    assert np.array_equal(clip10[:64], GOLD["pcm_head"])
    assert abs(float(clip10.astype(np.float64).sum()) - float(GOLD["pcm_sum"])) < 1e-9


def test_golden_config0_clip(oracle, clip10):
    S = oracle.stft(clip10)
    assert S.shape == (1243, 512)
    assert np.allclose(S[GOLD["spec_rows"]], GOLD["spec_values"], rtol=0, atol=2e-6)
    assert abs(float(S.astype(np.float64).sum()) - float(GOLD["spec_sum"])) < 1e-2
    pk = oracle.peaks(S)
    assert np.array_equal(pk, GOLD["peaks"])
    h, t = oracle.hashes(pk)
    assert np.array_equal(h, GOLD["hash"]) and np.array_equal(t, GOLD["t_anchor"])


def test_window_table(oracle):
    assert np.array_equal(oracle.window(), npo.window())


@pytest.mark.parametrize("n", [0, 1000, 1023, 1024, 1151, 1152, 5000])
def test_stft_matches_numpy(oracle, n):
    x = np.random.default_rng(n).uniform(-1, 1, n).astype(np.float32)
    a, b = oracle.stft(x), npo.stft(x)
    assert a.shape == b.shape
    assert np.allclose(a, b, rtol=0, atol=4e-6)


def test_stft_pure_tone_lands_in_its_bin(oracle):
    t = np.arange(16000)
    S = oracle.stft(np.sin(2 * np.pi * 1000.0 * t / 16000).astype(np.float32))   # bin 64
    assert (S.argmax(axis=1) == 64).all()
    assert abs(S[0, 64] - np.log1p((0.54 * 1024 / 2) ** 2)) < 0.01


def test_peaks_match_scipy(oracle):
    rng = np.random.default_rng(1)
    for T in (1, 5, 24, 25, 26, 100, 300, 513):
        S = rng.gamma(2.0, 1.0, (T, 512)).astype(np.float32)
        assert np.array_equal(oracle.peaks(S), npo.peaks(S)), T
    S = oracle.stft(synth.make_track(3, 4.0))
    assert np.array_equal(oracle.peaks(S), npo.peaks(S))


def test_peaks_ties_and_gates(oracle):
    S = np.random.default_rng(5).uniform(0.01, 0.4, (60, 512)).astype(np.float32)
    S[:, 100] = 3.0
    S[10:13, 300:302] = 4.0
    S[30, 5] = 9.0
    pk = oracle.peaks(S)
    assert np.array_equal(pk, npo.peaks(S))
    t, f = pk >> 9, pk & 511
    assert ((f == 100).sum() >= 40) and ((t >= 10) & (t < 13) & (f >= 300) & (f < 302)).sum() == 6
    assert (f >= 9).all()
    assert len(oracle.peaks(np.zeros((40, 512), np.float32))) == 0          # silence: nothing above the gate


def test_peak_capacity_rules(oracle):
    with pytest.raises(OverflowError):
        oracle.peaks(np.ones((30, 512), np.float32))                          # > 64 row candidates
    S = np.random.default_rng(2).uniform(0.01, 0.4, (256, 512)).astype(np.float32)
    S[:, 20::56] = 5.0                                                        # 9 tied lines x 256 frames = 2304 > 2048
    with pytest.raises(OverflowError):
        oracle.peaks(S)
    S = np.random.default_rng(2).uniform(0.01, 0.4, (256, 512)).astype(np.float32)
    S[:, 20::64] = 5.0                                                        # 8 lines x 256 = 2048: exactly at capacity
    assert len(oracle.peaks(S)) == 2048


def test_hashes_match_python_loops(oracle):
    rng = np.random.default_rng(9)
    for n in (0, 1, 2, 40, 600):
        keys = np.unique((np.sort(rng.integers(0, 120, n)).astype(np.uint32) << 9) | rng.integers(9, 512, n).astype(np.uint32))
        a, b = oracle.hashes(keys.astype(np.uint32)), npo.hashes(keys.astype(np.uint32))
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    h, t = oracle.hashes(GOLD["peaks"])
    f1, f2, dt = h >> 15, (h >> 6) & 511, h & 63
    assert ((dt >= 2) & (dt <= 33)).all() and (np.abs(f1.astype(int) - f2.astype(int)) <= 128).all() and (f1 != f2).all()
    assert (np.bincount(np.unique(np.stack([t, f1]), axis=1, return_inverse=True)[1]) <= 8).all()   # fan-out cap


def test_index_order_and_match_against_python(oracle):
    tracks = [synth.make_track(50 + k, 8.0) for k in range(5)]
    fps = [oracle.fingerprint(x) for x in tracks]
    H = np.concatenate([h for h, _ in fps]); T = np.concatenate([t for _, t in fps])
    TR = np.concatenate([np.full(len(h), k, np.uint32) for k, (h, _) in enumerate(fps)])
    ix = oracle.Index(H, TR, T)
    key = (ix.hash.astype(np.uint64) << 40) | (ix.track.astype(np.uint64) << 20) | ix.t.astype(np.uint64)
    assert (np.diff(key.astype(np.int64)) > 0).all()                         # (hash, track, t) strictly increasing
    assert ix.bucket[0] == 0 and ix.bucket[-1] == len(H) and (np.diff(ix.bucket.astype(np.int64)) >= 0).all()
    for q in range(6):
        clip, start = synth.make_query(tracks[q % 5], q, 5.0, 20.0)
        qh, qt = oracle.fingerprint(clip[:56000])
        rows = ix.match(qh, qt)
        ref = npo.match(ix.hash, ix.track, ix.t, qh, qt)
        assert [tuple(int(r[n]) for n in ("count", "track", "offset", "q_first", "q_last")) for r in rows] == ref
        assert len(rows) and rows["track"][0] == q % 5 and abs(int(rows["offset"][0]) - round(start / 128)) <= 1
    tomb = np.zeros(5, np.uint8); tomb[2] = 1
    clip, _ = synth.make_query(tracks[2], 2, 5.0, 20.0)
    qh, qt = oracle.fingerprint(clip)
    assert all(r["track"] != 2 for r in ix.match(qh, qt, tomb))


def test_batch_entry_equals_single(oracle):
    clips = [synth.make_track(70 + k, 2.0 + k) for k in range(4)] + [np.zeros(100, np.float32)]
    pcm = np.concatenate(clips); off = np.concatenate([[0], np.cumsum([len(c) for c in clips])])
    h, t, hoff, nh, npk, used = oracle.fingerprint_batch(pcm, off, threads=2)
    assert used >= 1 and nh[-1] == 0
    for i, c in enumerate(clips):
        a, b = oracle.fingerprint(c)
        assert np.array_equal(h[hoff[i]:hoff[i + 1]], a) and np.array_equal(t[hoff[i]:hoff[i + 1]], b)


# ---------------------------------------------------------------- properties of the domain (size-independent)
def test_hop_aligned_excerpt_keeps_hashes_and_shifts_time(oracle):
    """An excerpt that starts on a hop boundary sees the same frames as the track, so every landmark whose peaks'
    neighbourhoods lie inside the excerpt reappears with the same hash and t shifted by the start frame."""
    x = synth.make_track(3, 12.0)
    k = 173                                                   # start frame
    h, t = oracle.fingerprint(x)
    he, te = oracle.fingerprint(x[k * 128:k * 128 + 6 * 16000])
    full = {(int(a), int(b)) for a, b in zip(h, t)}
    inside = [(int(a), int(b) + k) for a, b in zip(he, te) if 13 <= int(b) < oracle.num_frames(6 * 16000) - 13 - 33 - 12]
    assert len(inside) > 50
    assert sum(p in full for p in inside) >= 0.95 * len(inside)   # the rest: fan-out truncated differently at the cut


def test_self_identification_offset_is_the_start_frame(oracle):
    tracks = [synth.make_track(40 + k, 15.0) for k in range(6)]
    fps = [oracle.fingerprint(x) for x in tracks]
    H = np.concatenate([h for h, _ in fps]); T = np.concatenate([t for _, t in fps])
    TR = np.concatenate([np.full(len(h), k, np.uint32) for k, (h, _) in enumerate(fps)])
    ix = oracle.Index(H, TR, T)
    for k, start in ((0, 0), (2, 37), (5, 611)):
        q = tracks[k][start * 128:start * 128 + int(3.5 * 16000)]
        rows = ix.match(*oracle.fingerprint(q))
        assert len(rows) >= 1 and rows["track"][0] == k and rows["offset"][0] == start
        assert rows["count"][0] >= 20 and np.all(np.diff(rows["count"].astype(np.int64)) <= 0)
        assert rows["q_first"][0] <= rows["q_last"][0] < oracle.num_frames(len(q))


@pytest.mark.parametrize("n", [0, 1, 1023, 1024, 1151, 1152])
def test_short_inputs(oracle, n):
    """Fewer than 1024 samples give no frame; 1024..1151 give one; no padding anywhere."""
    x = np.linspace(-0.5, 0.5, n, dtype=np.float32)
    frames = oracle.num_frames(n)
    assert frames == (0 if n < 1024 else (n - 1024) // 128 + 1)
    h, t = oracle.fingerprint(x)
    assert len(h) == 0 and len(t) == 0
    if frames:
        assert oracle.stft(x).shape == (frames, 512)


def test_deleted_track_disappears_and_only_it(oracle):
    tracks = [synth.make_track(60 + k, 10.0) for k in range(4)]
    tracks.append(tracks[1].copy())
    fps = [oracle.fingerprint(x) for x in tracks]
    H = np.concatenate([h for h, _ in fps]); T = np.concatenate([t for _, t in fps])
    TR = np.concatenate([np.full(len(h), k, np.uint32) for k, (h, _) in enumerate(fps)])
    ix = oracle.Index(H, TR, T)
    q = oracle.fingerprint(tracks[1][16000:16000 + 56000])
    both = ix.match(*q)
    assert [int(x) for x in both["track"][:2]] == [1, 4] and both["count"][0] == both["count"][1]
    tomb = np.zeros(5, np.uint8); tomb[1] = 1
    one = ix.match(*q, tombstone=tomb)
    assert int(one["track"][0]) == 4 and 1 not in one["track"]
    assert np.array_equal(one, both[both["track"] != 1])


def test_f32_streaming_baseline_follows_the_specification(oracle):
    """oracle/aid_cpu_f32.c (the CPU baseline bench.py states): spectrogram within AID_SPEC_TOL of the double-precision
    checker, peaks exactly aid_oracle_peaks of its OWN spectrogram (streaming ring == full-spectrogram definition),
    hashes by the shared hasher; ragged batch, short and edge lengths included."""
    from audio_ident_b200 import synth
    clips = [synth.make_track(40 + k, s) for k, s in enumerate([12.0, 0.064, 0.07, 1.0, 2.1, 3.3, 0.0])]
    clips.append(np.zeros(5000, np.float32))                    # silence: no peaks
    clips.append(np.tile(np.r_[1.0, np.zeros(127)], 80).astype(np.float32))      # impulse train: tie-heavy rows
    for x in clips:
        S, F = oracle.stft(x), oracle.stft_f32(x)
        assert S.shape == F.shape
        if S.size:
            assert (np.abs(F - S) <= 1e-4 * np.maximum(np.abs(S), 1.0)).all()
        try:
            ref = oracle.peaks(F)
        except OverflowError:
            with pytest.raises(OverflowError):
                oracle.peaks_f32(x)
            continue
        assert np.array_equal(oracle.peaks_f32(x), ref)
    pcm = np.concatenate(clips[:7])
    off = np.concatenate([[0], np.cumsum([len(c) for c in clips[:7]])])
    h, t, hoff, nh, npk, used = oracle.fingerprint_batch(pcm, off, 2, f32=True)
    for i, x in enumerate(clips[:7]):
        rh, rt = oracle.hashes(oracle.peaks_f32(x))
        assert np.array_equal(h[hoff[i]:hoff[i + 1]], rh) and np.array_equal(t[hoff[i]:hoff[i + 1]], rt)
    # and against the checker end to end: same hashes unless a float near-tie peak flips
    h64, t64, hoff64, *_ = oracle.fingerprint_batch(pcm, off, 2)
    same = sum(np.array_equal(h[hoff[i]:hoff[i + 1]], h64[hoff64[i]:hoff64[i + 1]]) for i in range(7))
    assert same >= 6


def test_resampler_definition_matches_scipy_resample_poly():
    """the decimator's definition (oracle/np_oracle.py resample3, what csrc/resample.cu
    implements) is scipy.signal.resample_poly(x, 1, 3); the engine's float32 tap table is scipy's firwin design."""
    import ctypes as C
    from scipy.signal import resample_poly
    from audio_ident_b200 import _lib
    L = _lib.load()
    taps = np.zeros(61, np.float32)
    L.aid_resample_taps(taps.ctypes.data_as(C.POINTER(C.c_float)))
    ref = npo.resample_taps()
    assert np.abs(taps - ref).max() < 2e-8 and abs(float(taps.astype(np.float64).sum()) - 1.0) < 1e-6
    assert np.array_equal(taps, taps[::-1])                      # linear phase
    assert [L.aid_resample_out_len(n) for n in (0, 1, 2, 3, 4, 48000)] == [0, 1, 1, 1, 2, 16000]
    rng = np.random.default_rng(3)
    for n in (1, 2, 3, 59, 61, 1000, 48001):
        x = rng.uniform(-1, 1, n)
        y = npo.resample3(x)
        z = resample_poly(x, 1, 3)
        assert y.shape == z.shape
        assert np.abs(y - z).max() < 1e-6
