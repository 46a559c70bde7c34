import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _build_checkers():
    """Compile the CPU checkers (oracle restatement; the reference harness when
    /root/reference is mounted) and the CUDA library if they are stale."""
    import oracle_lib
    oracle_lib.build_oracle()
    from ufomap_b200 import build as b
    if os.path.exists("/usr/local/cuda/bin/nvcc"):
        b.build()
    yield
