"""Seeded inputs of the content-duplicate golden vectors (shared by make_dedup_golden.py, which runs the reference
on them, and tests/test_dedup_*.py, which replay them). Only the reference's OUTPUTS are stored in
dedup_contract.json, together with a SHA-256 of these inputs, so the fixture stays small."""
import functools
import hashlib
import json
import random
import uuid


def rand_fp(rng, n, signed):
    vals = [rng.getrandbits(32) for _ in range(n)]
    if signed:
        vals = [v - (1 << 32) if v >= (1 << 31) else v for v in vals]      # what `fpcalc -raw -signed` prints
    return vals


def mutate(rng, vals, flip_prob):
    out = []
    for v in vals:
        m = 0
        for b in range(32):
            if rng.random() < flip_prob:
                m |= 1 << b
        out.append(((v & 0xFFFFFFFF) ^ m))
    return out


def text(vals):
    return ",".join(str(v) for v in vals)


@functools.lru_cache(maxsize=None)
def similarity_pairs():
    """(fp1, fp2) texts: the reference's own test inputs (tests/test_audio_dedup.py:137-168), grammar corner cases,
    then seeded random pairs."""
    rng = random.Random(42)
    pairs = [("100,200,300,400,500", "100,200,300,400,500"), ("0,0,0,0", "-1,-1,-1,-1"), ("", ""), ("abc", "def"),
             ("100,200", "100,200,300,400,500,600,700,800"), ("100,200,300", "100,201,300"),
             ("5", ""), ("1,2,", "1,2"), (" 7 , 8", "7,8"), ("4294967295", "-1"), ("4294967296,1", "0,1"),
             ("-4294967297", "4294967295"), ("1_0,2", "10,2"), ("1.5,2", "1,2"), ("0x10", "16"),
             ("123456789012345678901234567890,5", "0,5")]
    for _ in range(60):
        n1 = rng.choice([1, 2, 3, 31, 32, 33, 100, 233, 948, 1500])
        a = rand_fp(rng, n1, rng.random() < 0.5)
        kind = rng.random()
        if kind < 0.4:
            b = mutate(rng, a, rng.choice([0.0, 0.01, 0.05, 0.2, 0.5]))
        elif kind < 0.7:
            cut = rng.randint(1, n1)
            b = mutate(rng, a[:cut], 0.03) + rand_fp(rng, rng.randint(0, 40), False)
        else:
            b = rand_fp(rng, rng.choice([1, 7, 64, 233, 1000]), rng.random() < 0.5)
        pairs.append((text(a), text(b)))
    return pairs


@functools.lru_cache(maxsize=None)
def scan_cases():
    """dicts {rows: [(uuid str, fp text | None, duration | None)], fingerprint, duration, threshold}."""
    rng = random.Random(43)
    ids = [str(uuid.UUID(int=rng.getrandbits(128), version=4)) for _ in range(400)]
    cases = []

    def add(rows, fp, duration, threshold):
        cases.append({"rows": rows, "fingerprint": fp, "duration": duration, "threshold": threshold})

    # the reference's own cases (tests/test_audio_dedup.py:171-246)
    add([(ids[0], "100,200,300,400", 10.0)], "100,200,300,400", 10.0, 0.85)
    add([(ids[0], "4294967295,4294967295,4294967295,4294967295", 10.0)], "0,0,0,0", 10.0, 0.85)
    add([], "100,200,300", 10.0, 0.85)
    add([(ids[1], "999,888,777,666", 10.0), (ids[0], "100,200,300,400", 10.0)], "100,200,300,400", 10.0, 0.85)
    # ties: the first of two identical rows wins; duration window edges; NULL columns; unparsable rows and queries
    add([(ids[2], "1,2,3", 10.0), (ids[3], "1,2,3", 10.0)], "1,2,3", 10.0, 0.85)
    add([(ids[2], "1,2,3", 9.0), (ids[3], "1,2,3", 11.0), (ids[4], "1,2,3", 8.999999), (ids[5], "1,2,3", 11.000001)], "1,2,7", 10.0, 0.5)
    add([(ids[2], "1,2,3", 8.99), (ids[3], "1,2,3", 11.01)], "1,2,3", 10.0, 0.85)
    add([(ids[2], None, 10.0), (ids[3], "1,2,3", None), (ids[4], "x,y", 10.0), (ids[5], "", 10.0), (ids[6], "1,2,2", 10.0)], "1,2,3", 10.0, 0.85)
    add([(ids[2], "1,2,3", 10.0)], "not a fingerprint", 10.0, 0.85)
    add([(ids[2], "1,2,3", 10.0)], "", 10.0, 0.0)
    add([(ids[2], "0,0", 10.0)], "-1,-1", 10.0, 0.0)          # similarity exactly 0.0 never becomes the best
    add([(ids[2], "1,2,3", 0.0)], "1,2,3", 0.0, 0.85)
    for c in range(40):
        n_rows = rng.choice([1, 5, 40, 300])
        dur = rng.uniform(5.0, 400.0)
        base = rand_fp(rng, max(1, int(dur * 7.8)), True)
        rows = []
        for r in range(n_rows):
            kind = rng.random()
            d = dur * rng.uniform(0.85, 1.15)
            if kind < 0.15:
                f = text(mutate(rng, base, rng.choice([0.0, 0.02, 0.06, 0.1])))
            elif kind < 0.25:
                f = text(mutate(rng, base[: rng.randint(1, len(base))], 0.02))
            elif kind < 0.3:
                f = None if rng.random() < 0.5 else "garbage"
            else:
                f = text(rand_fp(rng, max(1, int(d * 7.8)), rng.random() < 0.5))
            rows.append((ids[r], f, None if rng.random() < 0.03 else d))
        if c % 5 == 0 and n_rows > 3:                                   # exact duplicates of one row: a tie
            k = rng.randrange(n_rows - 1)
            rows[-1] = (ids[n_rows - 1], rows[k][1], rows[k][2])
        add(rows, text(mutate(rng, base, rng.choice([0.0, 0.01, 0.04]))), dur, rng.choice([0.85, 0.85, 0.5, 0.95, 0.0]))
    return cases


def digest(obj) -> str:
    return hashlib.sha256(json.dumps(obj, sort_keys=True, separators=(",", ":")).encode()).hexdigest()
