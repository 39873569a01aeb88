"""pytest configuration: the `gpu` marker, import paths, shared fixtures."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """The C-ABI library, (re)built in-tree by nvcc (cross-compiles without a GPU)."""
    from cudatracerlib_b200 import build
    build.build()
    from cudatracerlib_b200 import lib
    return lib()


@pytest.fixture(scope="session")
def orc():
    import oracle_binding as ob
    ob.oracle()
    return ob
