import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "timeout: per-test time limit (pytest-timeout)")


def pytest_collection_modifyitems(config, items):
    """A hung kernel must end as a failed test, not as a stalled box: every GPU test gets a
    generous time limit when pytest-timeout is installed (the whole GPU suite runs in < 1 min)."""
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for item in items:
        if item.get_closest_marker("gpu") and not item.get_closest_marker("timeout"):
            item.add_marker(pytest.mark.timeout(600, method="thread"))


@pytest.fixture(scope="session")
def cuda_lib():
    """Loads the in-tree CUDA library; GPU tests fail loudly if it is missing."""
    from cobaya_b200 import _cabi

    return _cabi.load()
