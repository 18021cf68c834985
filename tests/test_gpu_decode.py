"""the GPU half of the decode feed. The 48 kHz -> 16 kHz decimator against its oracle
(scipy.signal.resample_poly, restated in oracle/np_oracle.py), and the drop-in decode module with a stand-in ffmpeg."""
import asyncio
import os
import stat

import numpy as np
import pytest

from oracle import np_oracle as npo

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [0, 1, 2, 3, 30, 31, 61, 767, 768, 769, 48000 * 3 + 1])
def test_resampler_matches_oracle(engine, n):
    rng = np.random.default_rng(n)
    t = np.arange(n)
    x = (0.5 * np.sin(2 * np.pi * 1000 * t / 48000) + 0.3 * np.sin(2 * np.pi * 11000 * t / 48000) +
         0.1 * rng.standard_normal(n)).astype(np.float32)
    y = engine.resample_48k_to_16k(x)
    ref = npo.resample3(x)
    assert y.shape == ref.shape and y.dtype == np.float32
    if n:
        assert np.abs(y - ref).max() <= 2e-6                 # float32 accumulation of 61 products of |x| <= ~1
        from scipy.signal import resample_poly
        assert np.abs(y - resample_poly(x.astype(np.float64), 1, 3)).max() <= 3e-6


def test_resampler_keeps_the_band_and_rejects_aliases(engine):
    t = np.arange(48000 * 2)
    keep = np.sin(2 * np.pi * 3000 * t / 48000).astype(np.float32)        # below the new Nyquist: passes
    alias = np.sin(2 * np.pi * 15000 * t / 48000).astype(np.float32)      # would fold to 1 kHz: must be gone
    yk, ya = engine.resample_48k_to_16k(keep), engine.resample_48k_to_16k(alias)
    assert abs(np.sqrt(np.mean(yk[100:-100] ** 2)) - np.sqrt(0.5)) < 0.01
    assert np.sqrt(np.mean(ya[100:-100] ** 2)) < 0.01


def test_decode_dual_rate_with_one_ffmpeg_child(engine, tmp_path, monkeypatch):
    """The drop-in keeps the reference's return shape (decode.py:74-87) with ONE child process: a stand-in `ffmpeg`
    that counts its invocations and emits 48 kHz PCM; the 16 kHz stream is the GPU decimation of it, and a
    fingerprint of it identifies the track that was indexed from a native 16 kHz rendering."""
    from audio_ident_b200 import decode as dec
    from audio_ident_b200 import fingerprint as fp
    monkeypatch.setenv("OLAF_DB", str(tmp_path / "db"))
    fp.shutdown()
    n48 = 48000 * 8
    t = np.arange(n48) / 48000.0
    rng = np.random.default_rng(1)
    x48 = np.zeros(n48)
    for _ in range(120):                                      # tone bursts below 7 kHz
        c, f, d = rng.uniform(0, 8), np.exp(rng.uniform(np.log(200), np.log(6500))), rng.uniform(0.05, 0.3)
        x48 += rng.uniform(0.05, 0.4) * np.exp(-0.5 * ((t - c) / (d / 4)) ** 2) * np.sin(2 * np.pi * f * t)
    x48 = (0.9 * x48 / np.abs(x48).max()).astype(np.float32)
    pcm_file = tmp_path / "pcm48.raw"
    pcm_file.write_bytes(x48.tobytes())
    count = tmp_path / "count"
    fake = tmp_path / "ffmpeg"
    fake.write_text(f"#!/bin/sh\ncat > /dev/null\necho x >> {count}\ncase \"$*\" in *' 48000 '*) cat {pcm_file};; *) exit 3;; esac\n")
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", f"{tmp_path}:{os.environ['PATH']}")
    pcm16, pcm48 = asyncio.run(dec.decode_dual_rate(b"container-bytes"))
    assert count.read_text().count("x") == 1                  # the reference would have spawned two
    assert pcm48 == x48.tobytes()
    y = np.frombuffer(pcm16, "<f4")
    assert len(y) == n48 // 3 and np.abs(y - npo.resample3(x48)).max() <= 2e-6
    assert dec.pcm_duration_seconds(pcm16, 16000) == 8.0
    with pytest.raises(dec.AudioDecodeError, match="too long"):
        asyncio.run(dec.decode_and_validate(b"container-bytes", max_duration=5.0))
    with pytest.raises(dec.AudioDecodeError, match="Empty"):
        asyncio.run(dec.decode_dual_rate(b""))
    # the decimated stream is good enough for the path it feeds
    import uuid
    tid = uuid.uuid4()
    assert asyncio.run(fp.olaf_index_track(pcm16, tid))
    rows = asyncio.run(fp.olaf_query(pcm16[16000 * 4 * 2:16000 * 4 * 7]))
    assert rows and rows[0].reference_path == str(tid) and abs(rows[0].reference_start - rows[0].query_start - 2.0) < 0.02
    fp.shutdown()
