"""GPU parity for stages a4-a6: CUDA path through the C ABI vs the CPU oracle.

Tolerances: spectrogram |dS| <= 1e-4 * max(|S|, 1) (AID_SPEC_TOL, BASELINE.json north_star "within 1e-4
relative"); peaks and hashes bit-exact when both sides consume the same stored spectrogram / peaks;
end to end (GPU spectrogram vs oracle spectrogram) peak sets may differ only at float near-ties, which
the test counts and explains one by one.
"""
import numpy as np
import pytest

from audio_ident_b200 import synth
from audio_ident_b200.engine import ragged

pytestmark = pytest.mark.gpu

TOL = 1e-4


def spec_close(a, b):
    return np.abs(a - b) <= TOL * np.maximum(np.abs(b), 1.0)


def test_stft_matches_oracle_10s(engine, oracle, clip10):
    S = engine.stft(clip10, [0, len(clip10)])
    R = oracle.stft(clip10)
    assert S.shape == R.shape == (1243, 512)
    ok = spec_close(S, R)
    assert ok.all(), f"max err {np.abs(S - R).max()} at {np.argwhere(~ok)[:5]}"


@pytest.mark.parametrize("n", [0, 1, 1023, 1024, 1025, 1151, 1152, 1280, 9343, 9344, 20000])
def test_stft_edge_lengths(engine, oracle, n):
    rng = np.random.default_rng(n)
    x = rng.uniform(-1, 1, n).astype(np.float32)
    S = engine.stft(x, [0, n])
    R = oracle.stft(x)
    assert S.shape == R.shape
    assert spec_close(S, R).all()


def test_stft_ragged_batch_equals_single(engine, oracle):
    clips = [synth.make_track(k, s) for k, s in enumerate([0.05, 1.0, 3.7, 0.0, 8.3, 2.0])]
    pcm, off = ragged(clips)
    S = engine.stft(pcm, off)
    R = np.concatenate([oracle.stft(c) for c in clips])
    assert S.shape == R.shape
    assert spec_close(S, R).all()
    # batching must not change a single bit of the GPU's own result
    row = 0
    for c in clips:
        T = oracle.num_frames(len(c))
        if T:
            assert np.array_equal(engine.stft(c, [0, len(c)]), S[row:row + T])
        row += T


def test_stft_full_scale_and_silence(engine, oracle):
    n = 16000
    t = np.arange(n)
    for x in (np.zeros(n, np.float32), np.ones(n, np.float32), (0.999 * np.sin(2 * np.pi * 1000 * t / 16000)).astype(np.float32),
              np.where(t % 2 == 0, 1.0, -1.0).astype(np.float32)):
        assert spec_close(engine.stft(x, [0, n]), oracle.stft(x)).all()


def test_peaks_bit_exact_on_gpu_spectrogram(engine, oracle, clip10):
    S = engine.stft(clip10, [0, len(clip10)])
    pk, off, st = engine.peaks(S, [0, S.shape[0]])
    assert st[0] == 0
    ref = oracle.peaks(S)
    assert np.array_equal(pk, ref)
    assert len(ref) > 100


def test_peaks_ragged_blocks_and_edges(engine, oracle):
    # frame counts around the 256-frame block size and the 12-frame halo
    rng = np.random.default_rng(7)
    frames = [1, 2, 12, 13, 25, 26, 255, 256, 257, 268, 269, 511, 512, 513, 700]
    specs = [rng.gamma(2.0, 1.0, (T, 512)).astype(np.float32) for T in frames]
    off = np.concatenate([[0], np.cumsum(frames)])
    pk, poff, st = engine.peaks(np.concatenate(specs), off)
    assert (st == 0).all()
    for i, S in enumerate(specs):
        assert np.array_equal(pk[poff[i]:poff[i + 1]], oracle.peaks(S)), f"track {i} T={frames[i]}"


def test_peaks_exact_ties_are_all_peaks(engine, oracle):
    # plateaus: every member of an exact tie is a peak (aid_params.h), in frequency and in time
    S = np.random.default_rng(5).uniform(0.01, 0.4, (300, 512)).astype(np.float32)
    S[:, 100] = 3.0            # a stationary line: peak in every frame
    S[40:45, 300:303] = 4.0    # a 5 x 3 plateau
    S[200, 20] = 2.0
    S[200, 5] = 9.0            # below AID_PEAK_MIN_BIN: never a peak, but it still shadows its neighbourhood
    pk, poff, st = engine.peaks(S, [0, 300])
    ref = oracle.peaks(S)
    assert st[0] == 0 and np.array_equal(pk, ref)
    assert len(ref) >= 300


def test_peaks_capacity_rules_fail_the_track(engine, oracle):
    ok_spec = np.random.default_rng(3).gamma(2.0, 1.0, (64, 512)).astype(np.float32)
    bad = np.full((64, 512), 1.0, np.float32)          # every point ties: > 64 row candidates
    off = [0, 64, 128, 192]
    pk, poff, st = engine.peaks(np.concatenate([ok_spec, bad, ok_spec]), off)
    assert st[0] == 0 and st[2] == 0 and st[1] & 1
    with pytest.raises(OverflowError):
        oracle.peaks(bad)
    ref = oracle.peaks(ok_spec)
    assert np.array_equal(pk[poff[0]:poff[1]], ref) and np.array_equal(pk[poff[2]:poff[3]], ref)


def test_hashes_bit_exact_same_order(engine, oracle, clip10):
    ref_pk = oracle.peaks(oracle.stft(clip10))
    h, t, off = engine.hashes(ref_pk, [0, len(ref_pk)])
    rh, rt = oracle.hashes(ref_pk)
    assert np.array_equal(h, rh) and np.array_equal(t, rt)
    assert len(rh) > 100


def test_hashes_dense_synthetic_peaks(engine, oracle):
    # dense random constellations exercise the fan-out cap, dt/df limits and the >32-candidate loop
    rng = np.random.default_rng(11)
    tracks = []
    for n in (0, 1, 2, 50, 900, 5000):
        t = np.sort(rng.integers(0, 400, n)).astype(np.uint32)
        f = rng.integers(9, 512, n).astype(np.uint32)
        keys = np.unique((t << 9) | f)
        tracks.append(keys.astype(np.uint32))
    off = np.concatenate([[0], np.cumsum([len(k) for k in tracks])])
    h, t, hoff = engine.hashes(np.concatenate(tracks), off)
    for i, k in enumerate(tracks):
        rh, rt = oracle.hashes(k)
        assert np.array_equal(h[hoff[i]:hoff[i + 1]], rh) and np.array_equal(t[hoff[i]:hoff[i + 1]], rt), i


def explain_peak_diffs(oracle_mod, S_ref, S_gpu, pk_ref, pk_gpu):
    """Every peak present on one side only must be a near-tie (oracle/parity.py, shared with bench.py's parity leg)."""
    from oracle import parity
    return parity.explain_peak_diffs(S_ref, S_gpu, pk_ref, pk_gpu)


def test_end_to_end_fingerprint_10s(engine, oracle, clip10):
    """BASELINE.json configs[0]: one 10 s clip, hashes compared with the CPU path."""
    h, t, hoff, st = engine.fingerprint(clip10, [0, len(clip10)])
    assert st[0] == 0
    S_ref = oracle.stft(clip10)
    pk_ref = oracle.peaks(S_ref)
    rh, rt = oracle.hashes(pk_ref)
    S_gpu = engine.stft(clip10, [0, len(clip10)])
    pk_gpu, _, _ = engine.peaks(S_gpu, [0, S_gpu.shape[0]])
    n_tie = explain_peak_diffs(oracle, S_ref, S_gpu, pk_ref, pk_gpu)
    # the fused path equals its own stages bit for bit
    gh, gt = oracle.hashes(pk_gpu)
    assert np.array_equal(h, gh) and np.array_equal(t, gt)
    if n_tie == 0:
        assert np.array_equal(h, rh) and np.array_equal(t, rt)
    assert n_tie <= 2, f"{n_tie} tie peaks in a 10 s clip"


def test_end_to_end_batch_of_tracks(engine, oracle):
    clips = [synth.make_track(100 + k, 6.0 + k) for k in range(6)] + [np.zeros(0, np.float32), np.zeros(500, np.float32)]
    pcm, off = ragged(clips)
    h, t, hoff, st = engine.fingerprint(pcm, off)
    assert (st[:6] == 0).all() and (st[6:] == 4).all()
    total_ties = 0
    for i, c in enumerate(clips):
        S_ref = oracle.stft(c)
        if S_ref.shape[0] == 0:
            assert hoff[i] == hoff[i + 1]
            continue
        S_gpu = engine.stft(c, [0, len(c)])
        pk_ref = oracle.peaks(S_ref)
        pk_gpu, _, _ = engine.peaks(S_gpu, [0, S_gpu.shape[0]])
        total_ties += explain_peak_diffs(oracle, S_ref, S_gpu, pk_ref, pk_gpu)
        gh, gt = oracle.hashes(pk_gpu)
        assert np.array_equal(h[hoff[i]:hoff[i + 1]], gh) and np.array_equal(t[hoff[i]:hoff[i + 1]], gt)
    assert total_ties <= 4


def test_sub_batching_does_not_change_results(engine):
    clips = [synth.make_track(200 + k, 2.0 + 0.5 * k) for k in range(9)]
    pcm, off = ragged(clips)
    a = engine.fingerprint(pcm, off)
    engine.set_max_batch_frames(600)        # forces several sub-batches through both stream slots
    try:
        b = engine.fingerprint(pcm, off)
    finally:
        engine.set_max_batch_frames(8 * 1024 * 1024)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_device_resident_path_equals_host_path(engine):
    clips = [synth.make_track(300 + k, 3.0) for k in range(4)]
    pcm, off = ragged(clips)
    h, t, hoff, st = engine.fingerprint(pcm, off)
    d = engine.device_alloc(pcm.nbytes)
    try:
        engine.to_device(d, pcm)
        res = engine.fingerprint_dev(d, off)
        doff = engine.to_host(res.d_hash_off, len(off), np.uint32)
        dh = engine.to_host(res.d_hash, int(doff[-1]), np.uint32)
        dt = engine.to_host(res.d_t_anchor, int(doff[-1]), np.uint32)
    finally:
        engine.device_free(d)
    assert np.array_equal(doff.astype(np.int64), hoff) and np.array_equal(dh, h) and np.array_equal(dt, t)


# ---- kernel selection (aid_engine_set_kernels): every choice must give the same bits ----------------------------------

def _ragged_mix():
    rng = np.random.default_rng(77)
    clips = [synth.make_track(400 + k, s) for k, s in enumerate([0.07, 0.2, 1.0, 2.04, 3.5, 8.19, 8.2, 16.4, 33.0])]
    clips += [rng.uniform(-1, 1, n).astype(np.float32) for n in (1024, 1151, 1152, 1280, 4095, 9343, 9344)]
    clips += [np.zeros(0, np.float32), np.zeros(700, np.float32), np.zeros(5000, np.float32)]
    return ragged(clips)


def test_packed_stft_kernel_is_bit_identical_to_the_scalar_one(engine):
    """csrc/stft.cu: k_stft_packed (f32x2 instructions, the default) performs the scalar kernel's operations on every
    element in the same order, so the stored spectrogram must not differ in a single bit (odd and even frame counts,
    units of fewer than 64 frames, tracks shorter than one unit)."""
    pcm, off = _ragged_mix()
    try:
        engine.set_kernels(0, False)
        ref = engine.stft(pcm, off)
        for variant in (5, 7, 3, 1, 13, 16):
            engine.set_kernels(variant, False)
            got = engine.stft(pcm, off)
            assert ref.shape == got.shape
            assert np.array_equal(ref.view(np.uint32), got.view(np.uint32)), f"variant {variant}"
    finally:
        engine.set_kernels(5, True)


def test_peak_summary_path_gives_the_same_fingerprints(engine):
    """The STFT kernel's group maxima (peak_summary, default) feed the peak kernel instead of the spectrogram rows:
    hashes, anchors, offsets and status words must be identical to the row-streaming path and to the round-1 kernels,
    also across sub-batches."""
    pcm, off = _ragged_mix()
    try:
        engine.set_kernels(0, False)
        ref = engine.fingerprint(pcm, off)
        for variant, summary in ((5, False), (5, True), (7, True)):
            engine.set_kernels(variant, summary)
            got = engine.fingerprint(pcm, off)
            for x, y in zip(ref, got):
                assert np.array_equal(x, y), (variant, summary)
        engine.set_kernels(5, True)
        engine.set_max_batch_frames(700)
        got = engine.fingerprint(pcm, off)
        for x, y in zip(ref, got):
            assert np.array_equal(x, y)
    finally:
        engine.set_max_batch_frames(8 * 1024 * 1024)
        engine.set_kernels(5, True)


def test_peak_summary_path_with_exact_ties(engine, oracle):
    """Plateaus of exactly equal values (every member of a tie is a peak) through the summary path: a stepped chirp whose
    spectrogram the oracle also sees; peaks are compared on the GPU's own spectrogram."""
    x = synth.make_track(901, 12.0)
    x[16000:48000] = 0.0                                   # silence: S = 0 everywhere, below the gate
    x[64000:96000] = x[32000 + 64000:64000 + 64000]
    try:
        engine.set_kernels(5, True)
        h1 = engine.fingerprint(x, [0, len(x)])
        S = engine.stft(x, [0, len(x)])
        pk, _, _ = engine.peaks(S, [0, S.shape[0]])       # row-streaming peak kernel on the same spectrogram
        rh, rt = oracle.hashes(pk)
        assert np.array_equal(h1[0], rh) and np.array_equal(h1[1], rt)
    finally:
        engine.set_kernels(5, True)
