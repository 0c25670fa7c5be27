import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs the compiled reference oracle/_ref/libvtmref.so")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import bindings
    return bindings.oracle()


@pytest.fixture(scope="session")
def ref_lib():
    from oracle import bindings
    lib = bindings.ref()
    if lib is None:
        pytest.skip("oracle/_ref/libvtmref.so not built (needs /root/reference: make -f oracle/Makefile.ref)")
    return lib
