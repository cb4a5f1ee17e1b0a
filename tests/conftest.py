"""pytest configuration: `gpu` marker for tests that need a B200; everything else runs on CPU."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ctx():
    """One dctz_gpu context on cuda:0 for the whole session.  No fallback: if the library or the
    device is missing the GPU tests fail, they do not skip."""
    import dctz_b200

    c = dctz_b200.Context(0)
    yield c
    c.close()
