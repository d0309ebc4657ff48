import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The oracle's C library is test infrastructure: build it once per session."""
    from oracle import c_oracle
    c_oracle.build()
    # libscgr.so must already be in-tree (built by __graft_entry__.build()); build if nvcc is around
    from scgaussian_b200 import build as b
    try:
        b.build_library()
    except Exception as e:  # pragma: no cover
        print("libscgr build skipped:", e)
    yield
