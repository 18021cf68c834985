#!/usr/bin/env python
"""Generates tests/golden/reference_contract.json by running the REFERENCE's own Python
(/root/reference/audio-ident-service/app/...) on seeded inputs.

What the reference pins for the hot path is the boundary grammar and everything downstream of an OlafMatch row
(SURVEY.md section 8c): `_parse_olaf_output` (app/audio/fingerprint.py:273-350) and the exact lane's
window slicing / consensus / aggregation / confidence / ranking (app/search/exact.py:70-399). This script
imports those modules (sqlalchemy, the DB session and the ORM model are stubbed: they are not on the path)
and records their outputs; tests/test_boundary_contract.py replays the inputs through audio_ident_b200.

Run in the build container only (the GPU box has no /root/reference):  python tests/golden/make_golden.py
"""
import asyncio
import json
import os
import random
import sys
import types
import uuid

REF = "/root/reference/audio-ident-service"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_contract.json")


def stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_reference():
    sys.path.insert(0, REF)
    stub("sqlalchemy", select=lambda *a, **k: None)
    stub("sqlalchemy.ext")
    stub("sqlalchemy.ext.asyncio", AsyncSession=object)
    stub("app.db")
    stub("app.db.session", async_session_factory=lambda: None)
    stub("app.models")
    stub("app.models.track", Track=object)
    import app.audio.fingerprint as fp
    import app.search.exact as ex
    return fp, ex


def row_dict(m):
    return {"match_count": m.match_count, "query_start": m.query_start, "query_stop": m.query_stop,
            "reference_path": m.reference_path, "reference_id": m.reference_id,
            "reference_start": m.reference_start, "reference_stop": m.reference_stop}


def main():
    fp, ex = load_reference()
    rng = random.Random(42)
    ids = [str(uuid.UUID(int=rng.getrandbits(128), version=4)) for _ in range(8)]
    gold = {"source": "MacPhobos/audio-ident audio-ident-service/app/audio/fingerprint.py + app/search/exact.py",
            "constants": {k: getattr(ex, k) for k in ("MIN_ALIGNED_HASHES", "STRONG_MATCH_HASHES",
                                                      "SHORT_CLIP_THRESHOLD_SEC", "SUB_WINDOWS", "SAMPLE_RATE")}}

    # ---- CSV grammar
    texts = [
        "42, 0.5, 3.2, my-track, 1001, 10.0, 12.7",
        "5, 0.0, 1.0, a, 1, 0.0, 1.0\n99, 0.0, 1.0, b, 2, 0.0, 1.0\n20, 0.0, 1.0, c, 3, 0.0, 1.0\n",
        "7; 0.1; 0.9; semi; 4; 1.5; 2.5",
        "match count, query start, query stop, path, id, ref start, ref stop\n12, 1, 2, x, 9, 3, 4",
        "garbage line\n\n  \n3, 0.0, 1.0, ok, 5, 0.0, 1.0\nx, 0.0, 1.0, bad, 5, 0.0, 1.0\n1, 2, 3",
        "8, 0.25, 3.0, " + ids[0] + ", 77, 100.125, 102.875, extra, fields",
        "",
        "10,1.0,2.0,tight,3,4.0,5.0",
        "4, 0.0, 1.0, f, 1.5, 0.0, 1.0",
    ]
    for _ in range(6):
        lines = []
        for _ in range(rng.randint(1, 12)):
            lines.append(f"{rng.randint(1, 200)}, {rng.uniform(0, 3):.3f}, {rng.uniform(3, 5):.3f}, {rng.choice(ids)}, "
                         f"{rng.randint(0, 10**6)}, {rng.uniform(0, 600):.3f}, {rng.uniform(600, 700):.3f}")
        texts.append("\n".join(lines))
    gold["parse"] = [{"stdout": t, "rows": [row_dict(m) for m in fp._parse_olaf_output(t)]} for t in texts]

    # ---- window slicing (exact.py:374-399) and duration (:361-371)
    win = []
    for n in (0, 1, 15999, 16000, 48000, 56000, 56001, 68000, 79999, 80000, 80001, 160000):
        pcm = bytes(4 * n)
        for a, b in [tuple(w) for w in ex.SUB_WINDOWS] + [(0.0, 0.0), (2.0, 1.0), (4.9, 9.0), (0.75, 4.250001)]:
            win.append({"n_samples": n, "start": a, "stop": b, "n_bytes": len(ex._extract_pcm_window(pcm, a, b)),
                        "duration": ex._pcm_duration_sec(pcm)})
    gold["windows"] = win

    # ---- confidence (exact.py:340-353)
    gold["confidence"] = [[n, ex._normalize_confidence(n)] for n in list(range(-2, 45)) + [100, 1000]]

    # ---- consensus (exact.py:220-293) and full-clip aggregation (:296-332)
    def rand_rows(k):
        rows = []
        for _ in range(k):
            name = rng.choice(ids + ["not-a-uuid", " " + ids[1] + " "])
            rows.append(fp.OlafMatch(rng.randint(1, 60), rng.uniform(0, 1), rng.uniform(2, 3.5), name,
                                     rng.randint(0, 999), round(rng.uniform(0, 300), 3), round(rng.uniform(300, 400), 3)))
        return rows

    def cand_list(cs):
        return [{"track": str(c.track_uuid), "aligned_hashes": c.aligned_hashes, "offset_seconds": c.offset_seconds} for c in cs]

    cons = []
    for _ in range(40):
        wins = [rand_rows(rng.choice([0, 0, 1, 2, 3, 5])) for _ in range(3)]
        cons.append({"windows": [[row_dict(m) for m in w] for w in wins], "candidates": cand_list(ex._consensus_score(wins))})
    gold["consensus"] = cons
    agg = []
    for _ in range(30):
        rows = rand_rows(rng.randint(0, 8))
        agg.append({"rows": [row_dict(m) for m in rows], "candidates": cand_list(ex._matches_to_candidates(rows))})
    gold["aggregate"] = agg

    # ---- the whole lane above the metadata join, with olaf_query replaced by canned rows keyed on the clip length
    lane = []

    async def run_case(n_samples, canned, max_results):
        calls = []

        async def fake_query(pcm):
            calls.append(len(pcm) // 4)
            return list(canned.get(len(pcm) // 4, []))

        async def fake_enrich(top, session=None):
            return top

        ex.olaf_query = fake_query
        ex._enrich_with_metadata = fake_enrich
        res = await ex.run_exact_lane(bytes(4 * n_samples), max_results=max_results)
        return calls, res

    for n_samples in (0, 16000, 48000, 72000, 80000, 80001, 160000):
        for _ in range(6):
            lens = {56000, 52000, 40000, 36000, n_samples, max(0, n_samples - 12000), max(0, n_samples - 24000), 24000, 12000, 16000, 4000, 48000}
            canned = {ln: rand_rows(rng.choice([0, 1, 2, 4])) for ln in lens}
            mr = rng.choice([1, 3, 10])
            calls, res = asyncio.run(run_case(n_samples, canned, mr))
            lane.append({"n_samples": n_samples, "max_results": mr,
                         "canned": {str(k): [row_dict(m) for m in v] for k, v in canned.items()},
                         "query_lengths": calls,
                         "result": [{"track": str(c.track_uuid), "aligned_hashes": c.aligned_hashes,
                                     "confidence": c.confidence, "offset_seconds": c.offset_seconds} for c in res]})
    gold["lane"] = lane

    with open(OUT, "w") as f:
        json.dump(gold, f, separators=(",", ":"))
    print("wrote", OUT, {k: len(v) for k, v in gold.items() if isinstance(v, list)})


if __name__ == "__main__":
    main()
