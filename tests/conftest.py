import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure).  Built on demand."""
    import ugsm_oracle

    ugsm_oracle.lib()
    return ugsm_oracle


@pytest.fixture(scope="session")
def capi():
    from ug_stereomatcher_b200 import capi as c

    if not os.path.exists(c.LIB_PATH):
        c.build()
    c.lib()
    return c


@pytest.fixture(scope="session")
def gpu(capi):
    """A live handle on cuda:0.  No fallback: creation failure fails the test."""
    h = capi.Handle()
    yield h
    h.close()
