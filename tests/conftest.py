import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import solo_oracle
    solo_oracle.build()
    return solo_oracle


@pytest.fixture(scope="session")
def synth():
    from ann_solo_b200 import synth as s
    return s


@pytest.fixture(scope="session")
def engine():
    """One device handle for the whole GPU session; fails loudly when the extension or the GPU
    is missing (no CPU fallback exists)."""
    from ann_solo_b200.engine import SoloEngine
    eng = SoloEngine(0)
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def small_world(synth):
    """A small seeded library/query set shared by several tests (charge 2 only for speed)."""
    lib = synth.make_library(6000, seed=1, decoy_seed=2)
    per_charge = synth.split_by_charge(lib)
    queries = synth.make_queries(lib, 300, seed=3)
    return lib, per_charge, queries


def canon_pairs(pairs, n):
    """Peak assignments compared as canonically sorted pair lists (the order inside
    exact product ties is unspecified by std::sort)."""
    p = np.asarray(pairs[:n], np.int64).reshape(-1, 2)
    return p[np.lexsort((p[:, 1], p[:, 0]))]
