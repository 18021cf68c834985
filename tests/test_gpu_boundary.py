"""The drop-in module on a real engine: same call shapes and conventions as the reference's
app.audio.fingerprint, persistence in the index directory, batch API, exact lane."""
import asyncio
import uuid

import numpy as np
import pytest

from audio_ident_b200 import exact_lane, synth
from audio_ident_b200 import fingerprint as fp

pytestmark = pytest.mark.gpu


@pytest.fixture()
def index_dir(tmp_path, monkeypatch):
    monkeypatch.setenv("OLAF_DB", str(tmp_path / "olaf_db"))
    fp.shutdown()
    yield tmp_path / "olaf_db"
    fp.shutdown()


def test_index_query_delete_roundtrip(index_dir):
    tracks = [synth.make_track(400 + k, 12.0) for k in range(5)]
    ids = [uuid.uuid4() for _ in tracks]
    for x, tid in zip(tracks[:2], ids[:2]):
        assert asyncio.run(fp.olaf_index_track(x.tobytes(), tid)) is True
    assert asyncio.run(fp.index_tracks([(x.tobytes(), tid) for x, tid in zip(tracks[2:], ids[2:])])) == [True, True, True]
    clip, start = synth.make_query(tracks[3], 9, 5.0, 20.0)
    rows = asyncio.run(fp.olaf_query(clip.tobytes()))
    assert rows and isinstance(rows[0], fp.OlafMatch)
    assert rows[0].reference_path == str(ids[3]) and uuid.UUID(rows[0].reference_path) == ids[3]
    assert [r.match_count for r in rows] == sorted((r.match_count for r in rows), reverse=True)
    assert abs(rows[0].reference_start - rows[0].query_start - start / 16000) < 0.02
    assert rows[0].query_start <= rows[0].query_stop and rows[0].reference_start <= rows[0].reference_stop
    # the exact lane on top: three sub-windows in one engine call, consensus, threshold, confidence
    res = exact_lane.score_clips([clip.tobytes(), b"", tracks[1][:16000 * 8].tobytes()], max_results=3)
    assert res[0] and res[0][0].track_uuid == ids[3] and res[0][0].aligned_hashes >= 8 and 0 < res[0][0].confidence <= 1.0
    assert res[1] == []
    assert res[2] and res[2][0].track_uuid == ids[1] and abs(res[2][0].offset_seconds) < 0.2
    # the one-upload window form (offsets into the clip) equals the three separate window queries of the reference's loop
    res_bytes = exact_lane.score_clips([clip.tobytes(), b"", tracks[1][:16000 * 8].tobytes()], max_results=3,
                                       query_many=fp.query_many_sync)
    assert res == res_bytes
    wins = [(0, 0, 56000), (0, 12000, 68000), (0, 24000, 80000)]
    by_offsets = fp.query_windows_sync([clip.tobytes()], wins)
    by_copies = fp.query_many_sync([clip[a:b].tobytes() for _, a, b in wins])
    assert by_offsets == by_copies and all(by_offsets)
    # delete
    assert asyncio.run(fp.olaf_delete_track(ids[3])) is True
    assert asyncio.run(fp.olaf_delete_track(ids[3])) is False
    assert all(r.reference_path != str(ids[3]) for r in asyncio.run(fp.olaf_query(clip.tobytes())))


def test_index_survives_restart_and_wipe(index_dir):
    a, b = synth.make_track(500, 10.0), synth.make_track(501, 10.0)
    ia, ib = uuid.uuid4(), uuid.uuid4()
    assert asyncio.run(fp.olaf_index_track(a.tobytes(), ia))
    fp.checkpoint()                                   # snapshot
    assert asyncio.run(fp.olaf_index_track(b.tobytes(), ib))   # journal only
    assert asyncio.run(fp.olaf_delete_track(ia))
    before = asyncio.run(fp.olaf_query(b[16000:16000 * 7].tobytes()))
    fp.shutdown()                                     # "process restart"
    after = asyncio.run(fp.olaf_query(b[16000:16000 * 7].tobytes()))
    assert after == before and after[0].reference_path == str(ib)
    assert asyncio.run(fp.olaf_query(a[16000:16000 * 7].tobytes())) == []
    fp.shutdown()
    for f in index_dir.iterdir():                     # make rebuild-index: rm -rf $OLAF_LMDB_PATH/*
        f.unlink()
    assert asyncio.run(fp.olaf_query(b[16000:16000 * 7].tobytes())) == []


def test_conventions_for_bad_input(index_dir):
    tid = uuid.uuid4()
    assert asyncio.run(fp.olaf_index_track(b"", tid)) is False
    assert asyncio.run(fp.olaf_query(b"")) == []
    assert asyncio.run(fp.olaf_index_track(np.zeros(300, np.float32).tobytes(), tid)) is True    # too short: stored, no hashes
    assert asyncio.run(fp.olaf_query(np.zeros(300, np.float32).tobytes())) == []
    tie = np.tile(np.r_[1.0, np.zeros(127)], 400).astype(np.float32)           # impulse train: every frame identical, flat comb
    assert asyncio.run(fp.olaf_index_track(tie.tobytes(), uuid.uuid4())) in (True, False)         # never raises
    assert asyncio.run(fp.olaf_delete_track(uuid.uuid4())) is False


def test_queries_do_not_block_the_event_loop(index_dir):
    x = synth.make_track(600, 20.0)
    asyncio.run(fp.olaf_index_track(x.tobytes(), uuid.uuid4()))

    async def both():
        ticks = 0

        async def ticker():
            nonlocal ticks
            for _ in range(50):
                await asyncio.sleep(0.001)
                ticks += 1

        t = asyncio.create_task(ticker())
        res = await asyncio.gather(*[fp.olaf_query(x[16000 * k:16000 * (k + 5)].tobytes()) for k in range(4)])
        await t
        return ticks, res

    ticks, res = asyncio.run(both())
    assert ticks == 50 and all(r for r in res)


def test_concurrent_queries_share_engine_calls(index_dir):
    """64 olaf_query coroutines in flight at once (what the unmodified exact lane and
    concurrent requests produce, exact.py:150-171) are served by at most 4 engine calls, and every caller gets exactly
    the rows a sequential call returns."""
    tracks = [synth.make_track(700 + k, 12.0) for k in range(8)]
    ids = [uuid.uuid4() for _ in tracks]
    assert all(asyncio.run(fp.index_tracks([(x.tobytes(), tid) for x, tid in zip(tracks, ids)])))
    clips = []
    for q in range(64):
        clip, _ = synth.make_query(tracks[q % 8], 300 + q, 5.0, 20.0)
        clips.append(clip[12000 * (q % 3):12000 * (q % 3) + 56000].tobytes())       # the three sub-window positions
    sequential = [asyncio.run(fp.olaf_query(c)) for c in clips]
    assert sum(1 for r in sequential if r) >= 60

    async def burst():
        return await asyncio.gather(*[fp.olaf_query(c) for c in clips])

    calls0, windows0 = fp._batcher.engine_calls, fp._batcher.windows
    concurrent = asyncio.run(burst())
    assert fp._batcher.windows - windows0 == 64
    assert fp._batcher.engine_calls - calls0 <= 4
    assert concurrent == sequential
    # a failing engine call reaches every waiter as OlafError, not a hang
    real = fp.query_many_sync
    try:
        def boom(_clips):
            raise RuntimeError("injected")
        fp.query_many_sync = boom
        with pytest.raises(fp.OlafError):
            asyncio.run(fp.olaf_query(clips[0]))
    finally:
        fp.query_many_sync = real
    assert asyncio.run(fp.olaf_query(clips[0])) == sequential[0]


def test_damaged_journal_is_refused_not_uploaded(index_dir):
    """journal bytes go through the engine's range checks; replay stops at the damaged record and the
    records before it survive."""
    import struct
    a, b, c = (synth.make_track(800 + k, 8.0) for k in range(3))
    ia, ib, ic = uuid.uuid4(), uuid.uuid4(), uuid.uuid4()
    for x, tid in ((a, ia), (b, ib), (c, ic)):
        assert asyncio.run(fp.olaf_index_track(x.tobytes(), tid))
    fp.shutdown()
    jp = index_dir / "journal.bin"
    data = bytearray(jp.read_bytes())
    # second record: corrupt its first hash to a value outside the 24-bit hash space
    pos = 0
    magic, kind, name_len, n_frames, n_hash = struct.unpack_from("<4sIIqI", data, pos)
    pos += 24 + name_len + 8 * n_hash
    magic, kind, name_len, n_frames, n_hash = struct.unpack_from("<4sIIqI", data, pos)
    assert magic == b"AIDJ" and n_hash > 0
    struct.pack_into("<I", data, pos + 24 + name_len, 0xFFFFFFFF)
    jp.write_bytes(bytes(data))
    rows = asyncio.run(fp.olaf_query(a[16000:16000 * 6].tobytes()))
    assert rows and rows[0].reference_path == str(ia)                  # the record before the damage is there
    assert asyncio.run(fp.olaf_query(b[16000:16000 * 6].tobytes())) == []      # the damaged one and what follows are not
    assert asyncio.run(fp.olaf_query(c[16000:16000 * 6].tobytes())) == []


def test_snapshot_drops_tombstoned_postings(index_dir):
    """deletes are compacted away when a snapshot is written -- the postings of a deleted (or replaced)
    track do not travel from snapshot to snapshot; results are unchanged."""
    a, b = synth.make_track(900, 10.0), synth.make_track(901, 10.0)
    ia, ib = uuid.uuid4(), uuid.uuid4()
    assert asyncio.run(fp.index_tracks([(a.tobytes(), ia), (b.tobytes(), ib)])) == [True, True]
    both = fp.get_engine().index_stats()["postings"]
    assert asyncio.run(fp.olaf_delete_track(ia))
    before = asyncio.run(fp.olaf_query(b[16000:16000 * 7].tobytes()))
    fp.checkpoint()
    assert (index_dir / ".lock").exists() and not (index_dir / "journal.bin").exists()
    fp.shutdown()
    st = fp.get_engine().index_stats()
    assert 0 < st["postings"] < both and st["tracks"] == 1 and st["tracks_total"] == 2
    assert asyncio.run(fp.olaf_query(b[16000:16000 * 7].tobytes())) == before
    assert asyncio.run(fp.olaf_query(a[16000:16000 * 7].tobytes())) == []
