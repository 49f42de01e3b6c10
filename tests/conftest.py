import os
import sys

import pytest

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_small():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "ref_small.npz")))


@pytest.fixture(scope="session")
def golden_ref32():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "ref_ref32.npz")))
