import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def engine():
    """One engine for the whole GPU session. Fails loudly (no skip) when the CUDA path is unusable."""
    from audio_ident_b200.engine import Engine
    eng = Engine(0)
    yield eng
    eng.close()


@pytest.fixture()
def fresh_index(engine):
    engine.index_clear()
    yield engine
    engine.index_clear()


@pytest.fixture(scope="session")
def clip10():
    from audio_ident_b200 import synth
    return synth.make_track(0, 10.0)


def sorted_pairs(h, t):
    """canonical order for comparing (hash, t_anchor) sets"""
    k = (np.asarray(t, np.uint64) << np.uint64(32)) | np.asarray(h, np.uint64)
    return np.sort(k)
