import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.join(ROOT, "oracle") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ll():
    """The product package (hyphenated directory name -> importlib)."""
    return importlib.import_module("light-loam_b200")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure only)."""
    import orc_py
    orc_py.build()
    return orc_py
