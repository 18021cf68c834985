"""The C-ABI library loads on a CPU-only box and exports every symbol include/audio_ident_b200.h declares;
without a GPU the product path fails loudly instead of falling back. CPU only (no compute calls)."""
import os
import re
import subprocess

import pytest

from audio_ident_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "audio_ident_b200.h")).read()


def declared():
    return sorted(set(re.findall(r"\b(aid_[a-z0-9_]+)\s*\(", HEADER)))


def test_header_and_binding_list_agree():
    assert declared() == sorted(_lib.EXPORTED)


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    for name in declared():
        assert hasattr(L, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (aid_[a-z0-9_]+)", out))
    assert set(declared()) <= exported
    assert L.aid_abi_version() == 1
    assert L.aid_strerror(0) == b"ok" and L.aid_strerror(-3) == b"output buffer too small"
    assert L.aid_num_frames(1023) == 0 and L.aid_num_frames(1024) == 1 and L.aid_num_frames(480000) == 3743


def test_header_cites_the_reference_interface_it_replaces():
    for needle in ("fingerprint.py:117-125", ":185-193", ":239-246", ":79-84"):
        assert needle in HEADER


def test_no_gpu_means_loud_failure_not_fallback():
    L = _lib.load()
    if L.aid_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    from audio_ident_b200.engine import Engine
    with pytest.raises(_lib.EngineUnavailable):
        Engine(0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "audio_ident_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "libaid_oracle" not in src, f
