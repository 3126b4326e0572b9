import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))
for p_ in (ROOT, HERE):
    if p_ not in sys.path:
        sys.path.insert(0, p_)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    """The oracle is test infrastructure: make sure it is compiled (and the reference objects too, where
    /root/reference is mounted) before any test touches it."""
    from oracle import pyoracle
    if not os.path.exists(pyoracle.LIB_PATH) or (os.path.isdir("/root/reference/src") and not pyoracle.ref_available()):
        pyoracle.build(ref=True)
    yield
