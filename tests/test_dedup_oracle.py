"""CPU: the dedup oracle (oracle/dedup_oracle.c) and the host-side parsing against outputs of the REFERENCE's own
functions (tests/golden/dedup_contract.json, made by tests/golden/make_dedup_golden.py from
audio-ident-service/app/audio/dedup.py:127-222). SURVEY.md section 8(f)-4; parity pinned, bit for bit."""
import json
import os
import sys
import uuid

import numpy as np
import pytest

GOLD_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD_DIR)
import dedup_cases as dc  # noqa: E402

from audio_ident_b200.dedup import parse_fingerprint  # noqa: E402  (pure Python; opens no device)


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLD_DIR, "dedup_contract.json")) as f:
        g = json.load(f)
    assert g["inputs_sha256"]["similarity"] == dc.digest(dc.similarity_pairs()), "dedup_cases.py drifted from the fixture"
    assert g["inputs_sha256"]["check"] == dc.digest(dc.scan_cases()), "dedup_cases.py drifted from the fixture"
    return g


def test_parse_follows_the_reference_grammar():
    assert parse_fingerprint("1,2,3").tolist() == [1, 2, 3]
    assert parse_fingerprint("-1, 4294967296 ,1_0").tolist() == [0xFFFFFFFF, 0, 10]
    for bad in ("", "abc", "1,2,", "1.5", "0x10", None):
        assert parse_fingerprint(bad) is None


def test_oracle_similarity_equals_the_reference(gold, oracle):
    pairs = dc.similarity_pairs()
    assert len(pairs) == len(gold["similarity"]) == 76
    for (a, b), want in zip(pairs, gold["similarity"]):
        wa, wb = parse_fingerprint(a), parse_fingerprint(b)
        got = 0.0 if wa is None or wb is None else oracle.fp_similarity(wa, wb)
        assert got.hex() == want, (a[:40], b[:40])
    # the reference's own assertions (tests/test_audio_dedup.py:137-168)
    vals = [float.fromhex(v) for v in gold["similarity"][:6]]
    assert vals[0] == 1.0 and vals[1] < 0.1 and vals[2] == 0.0 and vals[3] == 0.0 and 0.0 < vals[4] < 1.0 and 0.9 < vals[5] < 1.0


def pack_rows(rows):
    """what ContentStore.add_many keeps: rows the WHERE clause / loop can match, unparsable fingerprints empty"""
    ids, chunks, offs, durs = [], [], [0], []
    for tid, fp, d in rows:
        if fp is None or d is None:
            continue
        w = parse_fingerprint(fp)
        w = np.empty(0, np.uint32) if w is None else w
        ids.append(tid); chunks.append(w); offs.append(offs[-1] + w.size); durs.append(d)
    words = np.concatenate(chunks) if chunks else np.empty(0, np.uint32)
    return ids, words, np.asarray(offs, np.int64), np.asarray(durs, np.float64)


def test_oracle_scan_equals_the_reference(gold, oracle):
    cases = dc.scan_cases()
    assert len(cases) == len(gold["check"]) == 52
    n_hit = 0
    for c, want in zip(cases, gold["check"]):
        ids, words, off, dur = pack_rows(c["rows"])
        q = parse_fingerprint(c["fingerprint"])
        if q is None or not ids:
            got_id, got_sim = None, 0.0
        else:
            rows, sims = oracle.dedup_scan(words, off, dur, q, np.asarray([0, q.size]), np.asarray([c["duration"] * 0.9]),
                                           np.asarray([c["duration"] * 1.1]))
            got_sim = float(sims[0])
            got_id = ids[int(rows[0])] if rows[0] >= 0 else None
        assert got_sim.hex() == want["best_similarity"]
        accepted = got_id if (got_sim >= c["threshold"] and got_id is not None) else None    # dedup.py:214
        assert accepted == want["expected"]
        n_hit += accepted is not None
    assert 10 < n_hit < 52          # both outcomes are exercised


def test_oracle_scan_is_thread_count_independent(oracle):
    rng = np.random.default_rng(7)
    n = 5000
    lens = rng.integers(0, 300, n)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    words = rng.integers(0, 2**32, off[-1], dtype=np.uint64).astype(np.uint32)
    dur = rng.uniform(20, 40, n)
    words[off[4000]:off[4001]] = 0; words[off[100]:off[100] + 0] = 0
    q = rng.integers(0, 2**32, 240, dtype=np.uint64).astype(np.uint32)
    # plant the same near-duplicate twice: the earlier row must win at any thread count
    for r in (1234, 4321):
        ln = off[r + 1] - off[r]
        words[off[r]:off[r + 1]] = np.resize(q, ln)
        dur[r] = 30.0
    lens_equal = (off[1235] - off[1234]) == (off[4322] - off[4321])
    a = oracle.dedup_scan(words, off, dur, q, [0, q.size], [27.0], [33.0], n_threads=1)
    for t in (2, 5, 16):
        b = oracle.dedup_scan(words, off, dur, q, [0, q.size], [27.0], [33.0], n_threads=t)
        assert a[0].tolist() == b[0].tolist() and a[1].tobytes() == b[1].tobytes()
    assert a[0][0] in (1234, 4321) and (not lens_equal or a[0][0] == 1234)
