import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # building the native libraries is part of the test session (nvcc cross-compiles without a GPU)
    import __graft_entry__ as g
    g.build_trb()
    g.build_oracle()


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle as O
    return O.load_oracle("det")


@pytest.fixture(scope="session")
def oracle_sys():
    from oracle import pyoracle as O
    return O.load_oracle("sys")


@pytest.fixture(scope="session")
def trb():
    from tray_rust_b200 import _ffi
    return _ffi.load_trb()
