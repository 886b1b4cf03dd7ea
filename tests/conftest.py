import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree C-ABI library (compiled by nvcc here, prebuilt on the GPU box)."""
    import __graft_entry__ as g
    g.build()
    from mld_b200 import _lib
    return _lib.lib()


def golden(name):
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", name))
