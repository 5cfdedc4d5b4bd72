import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    return oracle_lib.Oracle()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference library; tests depending on it are skipped when oracle/_ref was not built."""
    import oracle_lib
    ref = oracle_lib.Reference()
    if not ref.available:
        pytest.skip("oracle/_ref/libgatbref.so not built (needs /root/reference)")
    return ref
