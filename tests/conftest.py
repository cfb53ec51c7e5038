import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_ops():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "ops.npz")))


@pytest.fixture(scope="session")
def golden_stream():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "stream_small.npz")))


@pytest.fixture(scope="session")
def golden_threeview():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "threeview.npz")))


@pytest.fixture(scope="session")
def golden_linear():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "linear.npz")))
