import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    """The builder's CPU restatement (oracle/port), the checker of the GPU tests."""
    from oracle.portdrv import PortOracle
    return PortOracle(threads=4)


@pytest.fixture(scope="session")
def ref():
    """The reference's own sources compiled against the shims (oracle/_ref)."""
    from oracle import refdrv
    if not refdrv.available():
        pytest.skip("oracle/_ref/libgiraffe_ref.so not built (needs /root/reference)")
    return refdrv.RefOracle(threads=4)
