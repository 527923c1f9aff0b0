import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def port():
    from oracle.oracle_py import Port
    return Port()


@pytest.fixture(scope="session")
def ref():
    from oracle.oracle_py import Reference
    if not Reference.available():
        pytest.skip("oracle/_ref/libdmc_ref.so not built and /root/reference absent")
    return Reference()
