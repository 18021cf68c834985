"""ctypes binding of libaudioident_b200.so (the C ABI in include/audio_ident_b200.h).

There is no fallback: if the shared library is missing or no CUDA device can be opened the
import / Engine construction raises. Nothing in this package imports oracle/.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libaudioident_b200.so")

MATCH_ROW_DTYPE = np.dtype([("count", "<i4"), ("track", "<u4"), ("offset", "<i4"),
                            ("q_first", "<i4"), ("q_last", "<i4")])

TRACK_OK, TRACK_PEAK_OVERFLOW, TRACK_TOO_LONG, TRACK_EMPTY = 0, 1, 2, 4


class FpDeviceResult(C.Structure):
    _fields_ = [("d_hash", C.c_void_p), ("d_t_anchor", C.c_void_p), ("d_hash_off", C.c_void_p),
                ("d_peaks", C.c_void_p), ("d_peak_off", C.c_void_p), ("d_status", C.c_void_p),
                ("d_spec", C.c_void_p), ("total_frames", C.c_int64)]


class EngineUnavailable(RuntimeError):
    """The CUDA engine cannot be used (library not built, or no device)."""


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineUnavailable(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or ./build_lib.sh). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i64p, i32p, u32p, u8p, f32p = (C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                        C.POINTER(C.c_uint32), C.POINTER(C.c_uint8), C.POINTER(C.c_float))
    f64p = C.POINTER(C.c_double)
    sig = {
        "aid_abi_version": (C.c_int, []),
        "aid_strerror": (C.c_char_p, [C.c_int]),
        "aid_get_params": (None, [i32p]),
        "aid_device_count": (C.c_int, []),
        "aid_engine_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "aid_engine_destroy": (None, [vp]),
        "aid_last_error": (C.c_char_p, [vp]),
        "aid_launch_count": (C.c_int64, [vp]),
        "aid_engine_sync": (C.c_int, [vp]),
        "aid_engine_set_max_batch_frames": (C.c_int, [vp, C.c_int64]),
        "aid_engine_set_stage_timing": (C.c_int, [vp, C.c_int]),
        "aid_engine_set_kernels": (C.c_int, [vp, C.c_int, C.c_int]),
        "aid_engine_stage_times": (C.c_int, [vp, C.POINTER(C.c_double), i64p]),
        "aid_fingerprint_host": (C.c_int, [vp, vp, i64p, C.c_int, vp, vp, C.c_int64, i64p, i32p]),
        "aid_fingerprint_dev": (C.c_int, [vp, vp, i64p, C.c_int, C.POINTER(FpDeviceResult), vp]),
        "aid_stft_host": (C.c_int, [vp, vp, i64p, C.c_int, vp]),
        "aid_peaks_host": (C.c_int, [vp, vp, i64p, C.c_int, vp, C.c_int64, i64p, i32p]),
        "aid_hashes_host": (C.c_int, [vp, vp, i64p, C.c_int, vp, vp, C.c_int64, i64p]),
        "aid_num_frames": (C.c_int64, [C.c_int64]),
        "aid_index_add_host": (C.c_int, [vp, vp, i64p, C.c_int, C.POINTER(C.c_char_p), u8p]),
        "aid_index_add_host_fp": (C.c_int, [vp, vp, i64p, C.c_int, C.POINTER(C.c_char_p), u8p, vp, vp, C.c_int64, i64p]),
        "aid_index_add_dev": (C.c_int, [vp, vp, i64p, C.c_int, C.POINTER(C.c_char_p), u8p]),
        "aid_index_add_hashes": (C.c_int, [vp, vp, vp, i64p, i64p, C.c_int, C.POINTER(C.c_char_p), u8p]),
        "aid_index_delete": (C.c_int, [vp, C.c_char_p]),
        "aid_index_commit": (C.c_int, [vp]),
        "aid_index_clear": (C.c_int, [vp]),
        "aid_index_set_grouping": (C.c_int, [vp, C.c_int]),
        "aid_index_stats": (C.c_int, [vp, i64p]),
        "aid_index_track_name": (C.c_int, [vp, C.c_uint32, C.c_char_p, C.c_int]),
        "aid_index_save": (C.c_int, [vp, C.c_char_p]),
        "aid_index_load": (C.c_int, [vp, C.c_char_p]),
        "aid_query_host": (C.c_int, [vp, vp, i64p, C.c_int, vp, C.c_int, i32p]),
        "aid_query_dev": (C.c_int, [vp, vp, i64p, C.c_int, vp, C.c_int, i32p]),
        "aid_query_windows_host": (C.c_int, [vp, vp, i64p, i64p, C.c_int, vp, C.c_int, i32p]),
        "aid_query_hashes": (C.c_int, [vp, vp, vp, i64p, C.c_int, vp, C.c_int, i32p]),
        "aid_match_dev": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_int, vp, C.c_int, vp, vp]),
        "aid_copy_device": (C.c_int, [vp, vp, vp, C.c_int64, vp]),
        "aid_exchange_create": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int64, C.POINTER(vp)]),
        "aid_exchange_destroy": (None, [vp]),
        "aid_exchange_handle": (C.c_int, [vp, vp]),
        "aid_exchange_connect": (C.c_int, [vp, vp]),
        "aid_exchange_connect_local": (C.c_int, [vp, C.POINTER(vp)]),
        "aid_exchange_set_timeout_ms": (C.c_int, [vp, C.c_int64]),
        "aid_exchange_status": (C.c_int, [vp]),
        "aid_match_exchange_dev": (C.c_int, [vp, vp, vp, vp, vp, vp, vp, C.c_int, vp, C.c_int64, vp, C.c_int, vp, vp]),
        "aid_identify_exchange_dev": (C.c_int, [vp, vp, vp, i64p, C.c_int, vp, C.c_int64, vp, C.c_int, vp, vp]),
        "aid_identify_exchange_host": (C.c_int, [vp, vp, vp, i64p, C.c_int, vp, C.c_int64, C.c_int, C.c_int, vp, C.c_int, vp]),
        "aid_identify_exchange_windows_dev": (C.c_int, [vp, vp, vp, i64p, i64p, C.c_int, vp, C.c_int64, vp, C.c_int, vp, vp]),
        "aid_identify_exchange_windows_host": (C.c_int, [vp, vp, vp, i64p, i64p, C.c_int, vp, C.c_int64, C.c_int, C.c_int, vp, C.c_int, vp]),
        "aid_fingerprint_windows_dev": (C.c_int, [vp, vp, i64p, i64p, C.c_int, C.POINTER(FpDeviceResult), vp]),
        "aid_synth_tracks_strided_dev": (C.c_int, [vp, vp, C.c_int64, C.c_int64, C.c_int, C.c_int64, C.c_uint64, vp]),
        "aid_match_stats": (C.c_int, [vp, i64p]),
        "aid_resample_out_len": (C.c_int64, [C.c_int64]),
        "aid_resample_taps": (None, [f32p]),
        "aid_resample_48k_to_16k_host": (C.c_int, [vp, vp, C.c_int64, vp]),
        "aid_resample_48k_to_16k_dev": (C.c_int, [vp, vp, C.c_int64, vp, vp]),
        "aid_device_alloc": (C.c_int, [vp, C.c_int64, C.POINTER(vp)]),
        "aid_device_free": (C.c_int, [vp, vp]),
        "aid_copy_to_device": (C.c_int, [vp, vp, vp, C.c_int64]),
        "aid_copy_to_host": (C.c_int, [vp, vp, vp, C.c_int64]),
        "aid_synth_tracks_dev": (C.c_int, [vp, vp, C.c_int64, C.c_int, C.c_int64, C.c_uint64, vp]),
        "aid_dedup_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "aid_dedup_destroy": (None, [vp]),
        "aid_dedup_last_error": (C.c_char_p, [vp]),
        "aid_dedup_size": (C.c_int64, [vp]),
        "aid_dedup_launch_count": (C.c_int64, [vp]),
        "aid_dedup_add": (C.c_int, [vp, vp, i64p, f64p, C.c_int, i64p]),
        "aid_dedup_scan": (C.c_int, [vp, vp, i64p, f64p, f64p, C.c_int, i64p, f64p]),
        "aid_dedup_last_scan_ms": (C.c_double, [vp]),
    }
    missing = []
    for name, (res, args) in sig.items():
        try:
            fn = getattr(L, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    if missing:
        raise EngineUnavailable(f"{LIB_PATH} lacks symbols {missing}: rebuild it")
    if L.aid_abi_version() != 1:
        raise EngineUnavailable("ABI version mismatch: rebuild the library")
    _lib = L
    return L


EXPORTED = [  # every symbol include/audio_ident_b200.h declares (tests/test_abi.py checks the header against this)
    "aid_abi_version", "aid_strerror", "aid_get_params", "aid_device_count", "aid_engine_create",
    "aid_engine_destroy", "aid_last_error", "aid_launch_count", "aid_engine_sync",
    "aid_engine_set_max_batch_frames", "aid_engine_set_stage_timing", "aid_engine_set_kernels", "aid_engine_stage_times", "aid_fingerprint_host", "aid_fingerprint_dev", "aid_stft_host",
    "aid_peaks_host", "aid_hashes_host", "aid_num_frames", "aid_index_add_host", "aid_index_add_host_fp", "aid_index_add_dev",
    "aid_index_add_hashes", "aid_index_delete", "aid_index_commit", "aid_index_clear", "aid_index_set_grouping", "aid_index_stats",
    "aid_index_track_name", "aid_index_save", "aid_index_load", "aid_query_host", "aid_query_dev", "aid_query_windows_host",
    "aid_query_hashes", "aid_match_dev", "aid_exchange_create", "aid_exchange_destroy", "aid_exchange_handle",
    "aid_exchange_connect", "aid_exchange_connect_local", "aid_exchange_set_timeout_ms", "aid_exchange_status",
    "aid_match_exchange_dev", "aid_identify_exchange_dev", "aid_identify_exchange_host", "aid_identify_exchange_windows_dev", "aid_identify_exchange_windows_host",
    "aid_fingerprint_windows_dev", "aid_synth_tracks_strided_dev", "aid_match_stats", "aid_resample_out_len", "aid_resample_taps",
    "aid_resample_48k_to_16k_host", "aid_resample_48k_to_16k_dev", "aid_copy_device", "aid_device_alloc", "aid_device_free", "aid_copy_to_device", "aid_copy_to_host",
    "aid_synth_tracks_dev", "aid_dedup_create", "aid_dedup_destroy", "aid_dedup_last_error", "aid_dedup_size",
    "aid_dedup_launch_count", "aid_dedup_add", "aid_dedup_scan", "aid_dedup_last_scan_ms",
]
