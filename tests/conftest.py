import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def infra():
    import ltelib
    ltelib.build_infra()
    return ltelib


@pytest.fixture(scope="session")
def phylib():
    """The product library; GPU tests fail (not skip) if it cannot be loaded."""
    import ltesniffer_b200
    return ltesniffer_b200.load_library()
