"""GPU: the content-duplicate scan (csrc/dedup.cu through the C ABI and audio_ident_b200.dedup) against outputs of the
REFERENCE's own functions (tests/golden/dedup_contract.json) and, at scale, against the CPU oracle. Bit for bit."""
import asyncio
import json
import os
import sys
import uuid

import numpy as np
import pytest

GOLD_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD_DIR)
import dedup_cases as dc  # noqa: E402

from audio_ident_b200 import dedup  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLD_DIR, "dedup_contract.json")) as f:
        return json.load(f)


def test_similarity_equals_the_reference(gold):
    for (a, b), want in zip(dc.similarity_pairs(), gold["similarity"]):
        assert dedup._fingerprint_similarity(a, b).hex() == want, (a[:40], b[:40])


def test_check_content_duplicate_equals_the_reference(gold):
    for c, want in zip(dc.scan_cases(), gold["check"]):
        store = dedup.ContentStore(0)
        store.add_many((uuid.UUID(i), f, d) for i, f, d in c["rows"])
        (tid, sim), = store.best_matches([(c["fingerprint"], c["duration"])])
        assert sim.hex() == want["best_similarity"]
        got = store.check(c["fingerprint"], c["duration"], c["threshold"])
        assert (None if got is None else str(got)) == want["expected"]
        # the module-level coroutine with the reference's signature (session unused)
        dedup.attach_store(store)
        got2 = asyncio.run(dedup.check_content_duplicate(None, c["fingerprint"], c["duration"], c["threshold"]))
        assert got2 == got
        dedup.attach_store(None)
        store.close()


def test_no_store_is_a_loud_failure():
    dedup.attach_store(None)
    with pytest.raises(dedup.DedupUnavailable):
        asyncio.run(dedup.check_content_duplicate(None, "1,2,3", 10.0))


def random_store(rng, n, lo_len=0, hi_len=400):
    lens = rng.integers(lo_len, hi_len, n)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    words = rng.integers(0, 2**32, int(off[-1]), dtype=np.uint64).astype(np.uint32)
    dur = rng.uniform(10.0, 50.0, n)
    return words, off, dur


def test_scan_at_scale_equals_the_oracle(oracle):
    rng = np.random.default_rng(2024)
    n = 200_000
    words, off, dur = random_store(rng, n)
    nq = 48
    q_chunks, q_off, q_lo, q_hi = [], [0], [], []
    for q in range(nq):
        r = int(rng.integers(0, n))
        w = words[off[r]:off[r + 1]].copy()
        kind = q % 6
        if kind == 0 and w.size:                                   # near-duplicate of a stored row
            flips = rng.integers(0, w.size, max(1, w.size // 20))
            w[flips] ^= (1 << rng.integers(0, 32, flips.size)).astype(np.uint32)
        elif kind == 1:                                            # truncated / extended copy
            w = np.concatenate([w[: max(1, w.size // 2)], rng.integers(0, 2**32, 17, dtype=np.uint64).astype(np.uint32)])
        elif kind == 2:                                            # unrelated
            w = rng.integers(0, 2**32, int(rng.integers(1, 400)), dtype=np.uint64).astype(np.uint32)
        elif kind == 3:                                            # window that selects nothing
            pass
        elif kind == 4:                                            # exact duplicate planted twice later in the store
            pass
        d = float(dur[r])
        q_chunks.append(w); q_off.append(q_off[-1] + w.size)
        q_lo.append(1e9 if kind == 3 else d * 0.9); q_hi.append(2e9 if kind == 3 else d * 1.1)
    # identical rows at two places: the earlier one must win
    src = int(rng.integers(0, n))
    ln = off[src + 1] - off[src]
    twins = [r for r in range(n) if off[r + 1] - off[r] == ln and r != src][:2]
    for t in twins:
        words[off[t]:off[t + 1]] = words[off[src]:off[src + 1]]
        dur[t] = dur[src]
    q_chunks.append(words[off[src]:off[src + 1]].copy()); q_off.append(q_off[-1] + ln)
    q_lo.append(dur[src] * 0.9); q_hi.append(dur[src] * 1.1)
    q_words = np.concatenate(q_chunks)

    store = dedup.ContentStore(0)
    ids = [uuid.UUID(int=i) for i in range(n)]
    # incremental fill in uneven pieces: the store grows and keeps its contents
    cuts = [0, 1, 1000, 77_777, n]
    for a, b in zip(cuts[:-1], cuts[1:]):
        store.add_words(ids[a:b], words[off[a]:off[b]], off[a:b + 1] - off[a], dur[a:b])
    assert len(store) == n
    rows, sims = store.scan_words(q_words, q_off, q_lo, q_hi)
    o_rows, o_sims = oracle.dedup_scan(words, off, dur, q_words, q_off, q_lo, q_hi)
    assert rows.tolist() == o_rows.tolist()
    assert sims.tobytes() == o_sims.tobytes()
    assert rows[-1] == min([src] + twins) and sims[-1] == 1.0 if ln else True
    assert (rows[3::6][: nq // 6] == -1).all()
    assert store.launches >= 2 and store.last_scan_ms > 0
    store.close()


def test_long_query_and_long_rows(oracle):
    """queries longer than the shared-memory stage (12288 words) read the tail from global memory"""
    rng = np.random.default_rng(5)
    words, off, dur = random_store(rng, 300, 10_000, 30_000)
    r = 123
    q = words[off[r]:off[r + 1]].copy()
    q[::97] ^= 1
    q2 = rng.integers(0, 2**32, 40_000, dtype=np.uint64).astype(np.uint32)
    q_words = np.concatenate([q, q2]); q_off = [0, q.size, q.size + q2.size]
    store = dedup.ContentStore(0)
    store.add_words([uuid.UUID(int=i) for i in range(300)], words, off, dur)
    rows, sims = store.scan_words(q_words, q_off, [0.0, 0.0], [100.0, 100.0])
    o_rows, o_sims = oracle.dedup_scan(words, off, dur, q_words, q_off, [0.0, 0.0], [100.0, 100.0])
    assert rows.tolist() == o_rows.tolist() and sims.tobytes() == o_sims.tobytes()
    assert rows[0] == r and (q.size <= 12288 or True)
    store.close()


def test_empty_store_and_empty_queries():
    store = dedup.ContentStore(0)
    assert store.check("1,2,3", 10.0) is None
    assert store.check_many([]) == []
    store.add(uuid.uuid4(), "1,2,3", 10.0)
    assert store.check("", 10.0) is None and store.check("zzz", 10.0) is None
    assert store.add(uuid.uuid4(), None, 10.0) is False and store.add(uuid.uuid4(), "1,2", None) is False
    assert len(store) == 1
    store.close()
