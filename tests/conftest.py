import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


class Golden:
    """Lazy npz reader: golden("returns")["cfg2/reward"]."""

    def __init__(self):
        self._cache = {}

    def __call__(self, group):
        if group not in self._cache:
            self._cache[group] = np.load(os.path.join(GOLDEN, group + ".npz"))
        return self._cache[group]


@pytest.fixture(scope="session")
def golden():
    return Golden()
