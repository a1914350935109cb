import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ref():
    """oracle/_ref/libmcxref.so: the reference's own kernel source built for the host (test infrastructure)."""
    from oracle import loader
    if not loader.have_ref():
        pytest.skip("oracle/_ref/libmcxref.so not built (python oracle/build_ref.py needs /root/reference)")
    return loader.ref()


@pytest.fixture(scope="session")
def port():
    """oracle/libmcxoracle.so: the plain-C restatement of the photon-transport algorithm."""
    from oracle import loader
    if not loader.have_port():
        pytest.skip("oracle/libmcxoracle.so not built (make -C oracle)")
    return loader.port()


@pytest.fixture(scope="session")
def lib():
    """The CUDA engine through its C ABI; fails loudly if it is not built."""
    from mcxcl_b200 import abi
    return abi.load()
