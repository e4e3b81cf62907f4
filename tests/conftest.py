import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle

    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def mh():
    import multih_b200

    if not os.path.exists(multih_b200.library_path()):
        multih_b200.build_library()
    return multih_b200


@pytest.fixture(scope="session")
def gpu_ctx(mh):
    """A live context on cuda:0 — fails loudly (no skip, no fallback) if the CUDA library cannot run."""
    return mh.Context()
