import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.dirname(os.path.abspath(__file__)) not in sys.path:
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))  # tests/peaks_util.py


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """Reference-generated fixtures (tests/golden/make_golden.py)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"))


@pytest.fixture(scope="session")
def golden_more():
    """Per-section fixtures of tests/golden/make_golden_more.py: golden_more("stencils") -> npz."""
    cache = {}

    def get(section):
        if section not in cache:
            cache[section] = np.load(os.path.join(ROOT, "tests", "golden", f"reference_golden_{section}.npz"))
        return cache[section]
    return get


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    o.build()
    return o
