import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
PKG = "project---hybrid-vector-search-queries_b200"
# The planner sends jobs below 4x10^6 pairs to the direct scan (a sweep's fixed cost would dominate).  The parity tests
# want the tile kernels (K2/K3) exercised on small, edge-case inputs too, so they switch that rule off.
os.environ.setdefault("HVS_MIN_TILE_PAIRS", "0")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def check():
    from oracle import check as Ck
    return Ck


@pytest.fixture(scope="session")
def datagen():
    return importlib.import_module(PKG + ".datagen")


@pytest.fixture(scope="session")
def hvs():
    """The product: ctypes binding over the C-ABI library (include/hvs.h)."""
    return importlib.import_module(PKG)


def load_golden(name):
    return np.load(os.path.join(GOLD, name))
