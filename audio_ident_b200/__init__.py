"""B200-native fingerprint-and-match engine behind the audio-ident service's fingerprint interface.

``audio_ident_b200.fingerprint`` mirrors ``app.audio.fingerprint`` of the reference service (same names, same
return conventions); ``audio_ident_b200.engine.Engine`` is the batch API underneath; the arithmetic lives in
``csrc/`` (sm_100a CUDA) behind the C ABI of ``include/audio_ident_b200.h``.
"""
__version__ = "0.1.0"
