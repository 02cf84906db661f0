import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "reference_vectors.npz")
    return np.load(path, allow_pickle=False)


@pytest.fixture(scope="session")
def headline():
    """Golden vectors of the configurations the headline numbers are quoted on + addition mode (make_golden.headline)."""
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "headline_vectors.npz")
    return np.load(path, allow_pickle=False)
