import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def port():
    from oracle.bindings import Port
    return Port()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference (oracle/_ref/libria_ref.so); skipped where it was never built."""
    from oracle.bindings import Ref
    if not Ref.available() and not os.path.isdir("/root/reference"):
        pytest.skip("oracle/_ref/libria_ref.so not built and /root/reference absent")
    return Ref()


@pytest.fixture(scope="session")
def ria_lib():
    """libria_b200.so, built on demand (nvcc cross-compiles without a GPU)."""
    import ria_b200
    if not os.path.exists(ria_b200.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return ria_b200.lib()


@pytest.fixture(scope="session")
def ctx(ria_lib):
    import torch
    import ria_b200
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    c = ria_b200.Context(0)
    yield c
    c.close()
