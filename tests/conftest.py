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


# ---- parity margins: every tolerance / arg-max comparison of the GPU tests records how far it was from failing.  The record is written
# at session end when MCAG_PARITY_MARGINS names a file (the GPU runs copy it to profiles/parity_margins.json).
_MARGINS = []


def record_margin(**kw):
    kw["test"] = os.environ.get("PYTEST_CURRENT_TEST", "").split(" ")[0]
    _MARGINS.append(kw)


def pytest_sessionfinish(session, exitstatus):
    path = os.environ.get("MCAG_PARITY_MARGINS")
    if path and _MARGINS:
        import json
        with open(path, "w") as f:
            json.dump({"tolerance": "|gpu - ref| <= 1e-6 + 1e-4 * max|ref| over the frame (ratio = err / that bound; element-wise relative error "
                                    "statistics beside it)", "records": _MARGINS}, f, indent=1)
