"""Long recordings: time-axis slices with a halo reproduce the single-pass fingerprint exactly; the window stream
finds the tracks a recording was assembled from (BASELINE.json configs[4], small scale)."""
import numpy as np
import pytest

from audio_ident_b200 import longform, sharded, synth
from audio_ident_b200.engine import ragged

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_slices", [2, 3, 7])
def test_sliced_fingerprint_equals_single_pass(engine, n_slices):
    pcm = np.concatenate([synth.make_track(950 + k, 9.0) for k in range(4)])       # 36 s, 4493 frames
    h, t, _, st = engine.fingerprint(pcm, [0, len(pcm)])
    assert st[0] == 0 and len(h) > 500
    hs, ts = longform.fingerprint_chunked(engine, pcm, n_slices)
    a = np.sort((t.astype(np.uint64) << np.uint64(32)) | h)
    b = np.sort((ts.astype(np.uint64) << np.uint64(32)) | hs)
    assert np.array_equal(a, b)
    assert np.array_equal(ts, np.sort(ts, kind="stable"))                          # slices come back in time order


def test_slice_plan_covers_everything():
    for T in (1, 255, 256, 257, 5000, 1349993):
        for n in (1, 2, 8):
            sl = longform.plan_slices(T, n)
            assert sl[0][0] == 0 and sl[-1][1] == T and all(x[1] == y[0] for x, y in zip(sl, sl[1:]))
            assert all(a % 256 == 0 for a, _ in sl)


def test_window_stream_recovers_the_playlist(engine):
    torch = pytest.importorskip("torch")
    tracks = [synth.make_track(970 + k, 20.0) for k in range(6)]
    engine.index_clear()
    sh = sharded.ShardedIdentifier(engine, 0, 1, device=torch.device("cuda", 0))
    pcm, off = ragged(tracks)
    assert sh.add(pcm, off, list(range(6))).all()
    rng = np.random.default_rng(4)
    order = [3, 0, 5, 1]
    gap = (rng.standard_normal(16000 * 6) * 0.01).astype(np.float32)
    rec = np.concatenate([np.concatenate([tracks[k], gap]) for k in order])
    rec = (rec + rng.standard_normal(len(rec)).astype(np.float32) * 0.01).astype(np.float32)
    segs, m, nn, starts = longform.identify_stream(sh, torch.from_numpy(rec).cuda(), len(rec), 10.0, 5.0, device=True)
    assert [s.track for s in segs] == order
    for s, k in zip(segs, range(len(order))):
        begin = k * 26.0
        assert abs(-s.offset_frames * 0.008 - begin) < 0.05                          # where the track starts in the recording
        assert s.start_s <= begin + 5.0 and s.stop_s >= begin + 15.0 and s.windows >= 2
    segs_h, *_ = longform.identify_stream(sh, rec, len(rec), 10.0, 5.0, device=False)
    assert [(s.track, s.offset_frames, s.votes) for s in segs_h] == [(s.track, s.offset_frames, s.votes) for s in segs]
    engine.index_clear()
