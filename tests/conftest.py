import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """Kernel parity first: under `-x` a failure in a CLI / pipeline test must not hide the kernel-parity suite
    (76 parity tests never ran on the driver's box behind one channel-noise failure)."""
    order = {"test_gpu_parity.py": 0, "test_fused_embed.py": 1, "test_configs.py": 2}
    items.sort(key=lambda it: order.get(os.path.basename(str(it.fspath)), 5))  # stable: the rest keeps its order


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
