import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    from oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def tool():
    """One workspace handle per test session; no CPU fallback -- fails loudly without a GPU."""
    import mh_spgemm_b200  # noqa: F401
    from mh_spgemm_b200 import api
    t = api.Tool(0)
    yield t
    t.release()
