import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle", "py")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    import orc as _orc
    _orc.build(ref=True)  # compiles the checker only; /root/reference is used when present (this container)
    return _orc


@pytest.fixture(scope="session")
def ref_available(orc):
    return orc.have_ref()
