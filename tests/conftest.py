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


@pytest.fixture(scope="session")
def golden_c1():
    """BASELINE config 1 (256 x 256 x 1536, the reference's laptop mode) run through the unmodified reference scripts:
    strided samples of every box, 24 sightlines, merged FLUX (tests/golden/run_reference_shimmed.py c1)."""
    import numpy as np
    path = os.path.join(GOLDEN, "ref_c1.npz")
    if not os.path.isfile(path):
        pytest.skip("tests/golden/ref_c1.npz not generated")
    return dict(np.load(path))
