"""What the reference pins for this path, replayed from golden vectors recorded from the
reference's own Python (tests/golden/make_golden.py -> reference_contract.json): the `olaf_c` CSV grammar
(app/audio/fingerprint.py:273-350) and the exact lane's slicing / consensus / ranking (app/search/exact.py).
CPU only."""
import json
import os
import uuid

import pytest

from audio_ident_b200 import exact_lane as xl
from audio_ident_b200 import fingerprint as fp

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_contract.json")))


def to_match(d):
    return fp.OlafMatch(d["match_count"], d["query_start"], d["query_stop"], d["reference_path"], d["reference_id"],
                        d["reference_start"], d["reference_stop"])


def test_public_names_and_dataclass_fields():
    import dataclasses
    import inspect
    for name in ("OlafError", "OlafMatch", "olaf_index_track", "olaf_query", "olaf_delete_track",
                 "_parse_olaf_output", "_parse_olaf_line", "_parts_to_match"):
        assert hasattr(fp, name), name
    assert [f.name for f in dataclasses.fields(fp.OlafMatch)] == [
        "match_count", "query_start", "query_stop", "reference_path", "reference_id", "reference_start", "reference_stop"]
    assert issubclass(fp.OlafError, Exception)
    for fn, params in ((fp.olaf_index_track, ["pcm_16k_f32le", "track_id"]), (fp.olaf_query, ["pcm_16k_f32le"]),
                       (fp.olaf_delete_track, ["track_id"])):
        assert inspect.iscoroutinefunction(fn)
        assert list(inspect.signature(fn).parameters) == params


def test_constants_match_reference():
    c = GOLD["constants"]
    assert xl.MIN_ALIGNED_HASHES == c["MIN_ALIGNED_HASHES"] and xl.STRONG_MATCH_HASHES == c["STRONG_MATCH_HASHES"]
    assert xl.SHORT_CLIP_THRESHOLD_SEC == c["SHORT_CLIP_THRESHOLD_SEC"] and xl.SAMPLE_RATE == c["SAMPLE_RATE"]
    assert [list(w) for w in xl.SUB_WINDOWS] == c["SUB_WINDOWS"]


@pytest.mark.parametrize("case", GOLD["parse"], ids=range(len(GOLD["parse"])))
def test_csv_grammar(case):
    got = fp._parse_olaf_output(case["stdout"])
    assert [m.__dict__ for m in got] == case["rows"]


def test_format_is_the_inverse_of_parse():
    m = fp.OlafMatch(42, 0.504, 3.2, str(uuid.uuid4()), 1001, 10.0, 12.704)
    assert fp._parse_olaf_line(fp.format_olaf_line(m)) == m


def test_window_slicing_and_duration():
    for w in GOLD["windows"]:
        pcm = bytes(4 * w["n_samples"])
        assert len(xl.extract_pcm_window(pcm, w["start"], w["stop"])) == w["n_bytes"], w
        assert xl.pcm_duration_sec(pcm) == w["duration"]


def test_confidence():
    for n, conf in GOLD["confidence"]:
        assert xl.normalize_confidence(n) == conf


def cands(cs):
    return [{"track": str(c.track_uuid), "aligned_hashes": c.aligned_hashes, "offset_seconds": c.offset_seconds} for c in cs]


def test_consensus_scoring():
    for case in GOLD["consensus"]:
        wins = [[to_match(d) for d in w] for w in case["windows"]]
        assert cands(xl.consensus_score(wins)) == case["candidates"]


def test_full_clip_aggregation():
    for case in GOLD["aggregate"]:
        assert cands(xl.matches_to_candidates([to_match(d) for d in case["rows"]])) == case["candidates"]


def test_whole_lane_with_canned_engine_rows():
    """Same engine calls (count, order, lengths) and the same ranked result as the reference's run_exact_lane."""
    for case in GOLD["lane"]:
        canned = {int(k): [to_match(d) for d in v] for k, v in case["canned"].items()}
        calls = []

        def fake_query_many(clips):
            calls.extend(len(c) // 4 for c in clips)
            return [list(canned.get(len(c) // 4, [])) for c in clips]

        res = xl.score_clips([bytes(4 * case["n_samples"])], case["max_results"], fake_query_many)[0]
        assert calls == case["query_lengths"], case["n_samples"]
        got = [{"track": str(c.track_uuid), "aligned_hashes": c.aligned_hashes, "confidence": c.confidence,
                "offset_seconds": c.offset_seconds} for c in res]
        assert got == case["result"]


def test_batched_lane_equals_clip_by_clip():
    cases = GOLD["lane"][:12]
    clips = [bytes(4 * c["n_samples"]) for c in cases]
    canned = {}
    for c in cases:
        for k, v in c["canned"].items():
            canned.setdefault(int(k), [to_match(d) for d in v])
    qm = lambda cl: [list(canned.get(len(x) // 4, [])) for x in cl]
    batched = xl.score_clips(clips, 10, qm)
    single = [xl.score_clips([c], 10, qm)[0] for c in clips]
    assert batched == single


def test_empty_inputs_do_not_touch_the_engine(monkeypatch):
    import asyncio

    def boom():
        raise AssertionError("engine must not be created for empty input")

    monkeypatch.setattr(fp, "get_engine", boom)
    assert asyncio.run(fp.olaf_index_track(b"", uuid.uuid4())) is False
    assert asyncio.run(fp.olaf_query(b"")) == []
    assert asyncio.run(fp.query_many([b"", b""])) == [[], []]
    assert asyncio.run(fp.index_tracks([(b"", uuid.uuid4())])) == [False]


def test_missing_engine_raises_olaf_error(monkeypatch):
    """Reference: missing binary -> OlafError (fingerprint.py:142-146). Here: no usable CUDA engine -> OlafError,
    never a silent CPU path."""
    import asyncio

    from audio_ident_b200 import _lib, engine

    def no_engine(*a, **k):
        raise _lib.EngineUnavailable("no CUDA device")

    fp.shutdown()
    monkeypatch.setattr(engine.Engine, "__init__", no_engine)
    with pytest.raises(fp.OlafError):
        asyncio.run(fp.olaf_query(bytes(64000)))
    with pytest.raises(fp.OlafError):
        asyncio.run(fp.olaf_index_track(bytes(64000), uuid.uuid4()))
    with pytest.raises(fp.OlafError):
        asyncio.run(fp.olaf_delete_track(uuid.uuid4()))


def test_window_ranges_equal_the_byte_slices():
    """exact_lane.plan_window_ranges (offsets into the clip, for the one-upload window query) describes exactly the
    byte strings plan_windows / the reference's _extract_pcm_window cut out (exact.py:374-399)."""
    import numpy as np
    from audio_ident_b200 import exact_lane as xl
    rng = np.random.default_rng(0)
    for n in (0, 1, 15999, 16000, 23999, 24000, 24001, 56000, 67999, 68000, 79999, 80000, 80001, 200000):
        pcm = rng.standard_normal(n).astype("<f4").tobytes()
        short, wins = xl.plan_windows(pcm)
        short2, ranges = xl.plan_window_ranges(pcm)
        assert short == short2 and len(wins) == len(ranges)
        for w, r in zip(wins, ranges):
            assert (pcm[r[0] * 4:r[1] * 4] if r is not None else b"") == w
