#!/usr/bin/env python
"""Generates tests/golden/dedup_contract.json by running the REFERENCE's own Python
(/root/reference/audio-ident-service/app/audio/dedup.py) on seeded inputs.

Pins SURVEY.md section 8(f)-4: `_fingerprint_similarity` (dedup.py:127-167) and `check_content_duplicate`
(dedup.py:170-222). sqlalchemy and the ORM model are stubbed (they are not on the path): the SQL WHERE clause of
dedup.py:192-197 is applied here with the same double comparisons before the rows are handed to the reference's
function through a mock session, exactly as the reference's own tests do (tests/test_audio_dedup.py:171-246).
Similarities are stored as C99 hex floats so the comparison is bit for bit.

Run in the build container only (the GPU box has no /root/reference):  python tests/golden/make_dedup_golden.py
"""
import asyncio
import json
import os
import sys
import types
import uuid

REF = "/root/reference/audio-ident-service"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dedup_contract.json")


class _Col:
    def isnot(self, other): return None
    def __ge__(self, other): return None
    def __le__(self, other): return None
    def __eq__(self, other): return None
    __hash__ = object.__hash__


class _Select:
    def where(self, *a): return self


def load_reference():
    sys.path.insert(0, REF)
    def stub(name, **attrs):
        m = types.ModuleType(name); m.__dict__.update(attrs); sys.modules[name] = m
    stub("sqlalchemy", select=lambda *a, **k: _Select())
    stub("sqlalchemy.ext")
    stub("sqlalchemy.ext.asyncio", AsyncSession=object)
    stub("app.models")
    track = type("Track", (), {k: _Col() for k in ("id", "file_hash_sha256", "chromaprint_fingerprint", "chromaprint_duration")})
    stub("app.models.track", Track=track)
    import app.audio.dedup as dd
    return dd


class _Result:
    def __init__(self, rows): self._rows = rows
    def all(self): return self._rows


class _Session:
    def __init__(self, rows): self._rows = rows
    async def execute(self, stmt): return _Result(self._rows)


def main():
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import dedup_cases as dc
    dd = load_reference()
    pairs, cases = dc.similarity_pairs(), dc.scan_cases()
    gold = {"source": "MacPhobos/audio-ident audio-ident-service/app/audio/dedup.py:127-222",
            "inputs_sha256": {"similarity": dc.digest(pairs), "check": dc.digest(cases)}}
    gold["similarity"] = [float(dd._fingerprint_similarity(a, b)).hex() for a, b in pairs]
    out = []
    for c in cases:
        # the WHERE clause of dedup.py:192-197, then the reference's function on the selected rows
        lo, hi = c["duration"] * 0.9, c["duration"] * 1.1
        selected = [(uuid.UUID(i), f, d) for i, f, d in c["rows"] if f is not None and d is not None and d >= lo and d <= hi]
        got = asyncio.run(dd.check_content_duplicate(_Session(selected), c["fingerprint"], c["duration"], c["threshold"]))
        sims = [float(dd._fingerprint_similarity(c["fingerprint"], f)) for _, f, _ in selected]
        out.append({"expected": None if got is None else str(got), "n_selected": len(selected),
                    "best_similarity": float(max(sims, default=0.0)).hex()})
    gold["check"] = out
    with open(OUT, "w") as f:
        json.dump(gold, f, indent=0, separators=(",", ":"))
    print(f"wrote {OUT}: {len(pairs)} similarity pairs, {len(cases)} scan cases, {os.path.getsize(OUT)} bytes")


if __name__ == "__main__":
    main()
