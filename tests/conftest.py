import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from tests import util
    return util.oracle()


@pytest.fixture(scope="session")
def reflib():
    from tests import util
    r = util.reflib()
    if r is None:
        pytest.skip("oracle/_ref/libmcmceq_ref.so not built (reference sources absent)")
    return r
