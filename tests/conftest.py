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
    config.addinivalue_line("markers", "timeout: per-test time limit (pytest-timeout)")


def pytest_collection_modifyitems(config, items):
    """A hung worker handshake or kernel must cost one test, not the GPU box's whole time budget: every test gets a
    limit (pytest-timeout, thread method so forked sampler workers do not outlive it silently)."""
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for item in items:
        if item.get_closest_marker("timeout") is None:
            item.add_marker(pytest.mark.timeout(150 if item.get_closest_marker("gpu") else 600))


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
