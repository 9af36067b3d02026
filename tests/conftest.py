import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "gkr-mimc_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The C oracle (oracle/gkr_oracle.c), built on demand. Test infrastructure only."""
    import coracle
    coracle.build()
    coracle.set_threads(min(8, os.cpu_count() or 1))
    return coracle


@pytest.fixture(scope="session")
def ctx():
    """One device context for the whole GPU session (max batch 2^16)."""
    import gkrb200
    c = gkrb200.Context(device=0, max_bn=16)
    yield c
    c.close()
