import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_points():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "pm_points.npz"))


@pytest.fixture(scope="session")
def golden_stages():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "pm_stages.npz"))


@pytest.fixture(scope="session")
def gpu_ctx():
    from sea_ice_drift_b200 import _lib
    ctx = _lib.Context(0)          # raises loudly when there is no CUDA device / library
    yield ctx
    ctx.close()
