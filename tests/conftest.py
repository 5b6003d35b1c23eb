import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")
    # a fresh checkout has no binaries: build libmidas_b200.so (nvcc cross-compiles without a GPU) and
    # the C oracle once, exactly as __graft_entry__.build() does
    lib_so = os.path.join(ROOT, "midas_b200", "libmidas_b200.so")
    oracle_so = os.path.join(ROOT, "oracle", "liboracle.so")
    if not (os.path.exists(lib_so) and os.path.exists(oracle_so)):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def ctx():
    """One libmidas_b200 context on cuda:0 for the whole GPU session (fails loudly without a GPU)."""
    from midas_b200 import lib
    c = lib.Context(0)
    yield c
    c.close()
