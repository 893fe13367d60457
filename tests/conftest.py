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
    from oracle import oracle_binding as ob
    return ob.Oracle()


@pytest.fixture(scope="session")
def surf():
    import megamol_b200 as mm
    s = mm.Surf(0)
    yield s
    s.close()


def rel_err(a, b, floor):
    """|a-b| / max(|b|, floor)"""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)
