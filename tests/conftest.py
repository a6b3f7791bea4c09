import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def port():
    from oracle import pyoracle
    return pyoracle.port()


@pytest.fixture(scope="session")
def ref():
    from oracle import pyoracle
    if not pyoracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return pyoracle.ref()


@pytest.fixture(scope="session")
def ctx():
    """The product path: CUDA library through the C ABI. Fails loudly if it cannot be created."""
    import steganosaurus_b200 as sb
    c = sb.Context(0)
    yield c
    c.close()
