import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def mnv():
    import mega_nerf_viewer_b200 as m

    return m


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py

    oracle_py.lib()  # builds liboracle.so on first use
    return oracle_py
