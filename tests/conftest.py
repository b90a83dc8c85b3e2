import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with -m gpu")


@pytest.fixture(scope="session")
def orc():
    import orclib
    return orclib.ORC()


@pytest.fixture(scope="session")
def ref():
    import orclib
    r = orclib.REF()
    if r is None:
        pytest.skip("oracle/_ref/libfmref.so not available on this host")
    return r


@pytest.fixture(scope="session")
def sdr():
    """The product binding; GPU tests fail (not skip) when the CUDA library is missing."""
    import sdr_b200
    sdr_b200.lib()
    return sdr_b200
