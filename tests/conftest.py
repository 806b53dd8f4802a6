"""pytest configuration: markers and import paths.

`-m "not gpu"`: oracle vs golden fixtures, host logic, C-ABI symbol checks (no GPU needed).
`-m gpu`      : parity tests proper; every one calls the CUDA kernels through the C ABI.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "hybrid-grid-for-hexagonal-and-rectangular-image-processing_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def resample_golden():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "resample_golden.npz"))


@pytest.fixture(scope="session")
def hexframes_golden():
    import numpy as np
    return np.load(os.path.join(GOLDEN, "hexframes_golden.npz"))
