import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle binding (test infrastructure; builds oracle/libklt_oracle.so on first use)."""
    from oracle import binding
    binding.lib()
    return binding


@pytest.fixture(scope="session")
def golden():
    import json
    import numpy as np
    gdir = os.path.join(ROOT, "tests", "golden")
    with open(os.path.join(gdir, "golden.json")) as f:
        meta = json.load(f)
    vec = dict(np.load(os.path.join(gdir, "solver_620x188.npz")))
    return meta, vec


@pytest.fixture(scope="session")
def tracker():
    """GPU tracker through the C ABI; fails loudly (no CPU fallback) if the library or GPU is missing."""
    from lego_slam_b200 import build
    build.build()
    import lego_slam_b200 as klt
    return klt.Tracker(0)
