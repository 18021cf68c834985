"""Drop-in for the reference's ``app.audio.decode`` (audio-ident-service/app/audio/decode.py) -- SURVEY.md section 8(f)-2.

Same names, signatures and error behaviour (``AudioDecodeError``, ``decode_to_pcm``, ``decode_dual_rate``,
``pcm_duration_seconds``, ``decode_and_validate``). What changes is ``decode_dual_rate``: the reference starts TWO ffmpeg
children per file, one per output rate (decode.py:74-87); here ONE child decodes to 48 kHz and the 16 kHz stream the
fingerprint path needs is derived on the GPU by a polyphase decimator (csrc/resample.cu, ``aid_resample_48k_to_16k_host``)
-- half the process spawns and half the decode work per ingested file and per search request.

The 16 kHz stream is the zero-phase 61-tap design of ``scipy.signal.resample_poly(x, 1, 3)`` (the oracle,
oracle/np_oracle.py ``resample3``); it is NOT bit-identical to what ``ffmpeg -ar 16000`` produces (libswresample has its
own filter) -- parity with the reference's decode is unpinned until an ffmpeg is available to record vectors from.
Set ``AID_DECODE_TWO_FFMPEG=1`` to get the reference's behaviour back. There is no CPU resampler here: without the CUDA
engine ``decode_dual_rate`` raises ``AudioDecodeError``.
"""
from __future__ import annotations

import asyncio
import logging
import os

logger = logging.getLogger(__name__)


class AudioDecodeError(Exception):
    """Raised when audio decoding fails."""


async def decode_to_pcm(audio_data: bytes, target_sample_rate: int, output_format: str = "f32le") -> bytes:
    """Decode audio to raw mono PCM with one ffmpeg child (reference decode.py:17-71): same command line, same errors."""
    if not audio_data:
        raise AudioDecodeError("Empty audio data provided")
    codec = f"pcm_{output_format}"
    proc = await asyncio.create_subprocess_exec(
        "ffmpeg", "-hide_banner", "-loglevel", "error", "-i", "pipe:0", "-ar", str(target_sample_rate), "-ac", "1",
        "-f", output_format, "-acodec", codec, "pipe:1",
        stdin=asyncio.subprocess.PIPE, stdout=asyncio.subprocess.PIPE, stderr=asyncio.subprocess.PIPE)
    stdout, stderr = await proc.communicate(input=audio_data)
    if proc.returncode != 0:
        err_msg = stderr.decode(errors="replace").strip()
        raise AudioDecodeError(f"ffmpeg exited with code {proc.returncode}: {err_msg}")
    if not stdout:
        raise AudioDecodeError("ffmpeg produced no output")
    return stdout


def resample_48k_to_16k_sync(pcm_48k_f32le: bytes) -> bytes:
    """48 kHz f32le mono -> 16 kHz f32le mono on the GPU (the engine of audio_ident_b200.fingerprint)."""
    from . import fingerprint
    try:
        eng = fingerprint.get_engine()
        with fingerprint._state.lock:
            return eng.resample_48k_to_16k(pcm_48k_f32le).tobytes()
    except Exception as exc:
        raise AudioDecodeError(f"GPU resampler not available ({exc}); there is no CPU fallback") from exc


async def decode_dual_rate(audio_data: bytes) -> tuple[bytes, bytes]:
    """``(pcm_16k_f32le, pcm_48k_f32le)`` as the reference returns them (decode.py:74-87), from ONE ffmpeg child."""
    if os.environ.get("AID_DECODE_TWO_FFMPEG") == "1":
        pcm_16k, pcm_48k = await asyncio.gather(
            decode_to_pcm(audio_data, target_sample_rate=16000, output_format="f32le"),
            decode_to_pcm(audio_data, target_sample_rate=48000, output_format="f32le"))
        return pcm_16k, pcm_48k
    pcm_48k = await decode_to_pcm(audio_data, target_sample_rate=48000, output_format="f32le")
    pcm_16k = await asyncio.get_running_loop().run_in_executor(None, resample_48k_to_16k_sync, pcm_48k)
    return pcm_16k, pcm_48k


def pcm_duration_seconds(pcm_data: bytes, sample_rate: int, sample_width: int = 4) -> float:
    """Duration from PCM bytes (reference decode.py:90-105)."""
    return len(pcm_data) / (sample_rate * sample_width)


async def decode_and_validate(audio_data: bytes, max_duration: float = 1800.0, min_duration: float = 0.0) -> tuple[bytes, bytes]:
    """Decode dual rate and validate duration constraints (reference decode.py:108-136): same messages."""
    pcm_16k, pcm_48k = await decode_dual_rate(audio_data)
    duration = pcm_duration_seconds(pcm_16k, sample_rate=16000, sample_width=4)
    if duration < min_duration:
        raise AudioDecodeError(f"Audio too short: {duration:.2f}s < minimum {min_duration:.2f}s")
    if duration > max_duration:
        raise AudioDecodeError(f"Audio too long: {duration:.2f}s > maximum {max_duration:.2f}s")
    return pcm_16k, pcm_48k
