import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """libstif_b200 context on cuda:0 (GPU tests only; fails loudly without a GPU)."""
    from stif_b200 import _native
    c = _native.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
