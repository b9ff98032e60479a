import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def csb():
    import cube_slam_wu_b200 as m
    from cube_slam_wu_b200 import build
    build.build()
    return m


@pytest.fixture(scope="session")
def ctx(csb):
    c = csb.Context(0)  # raises if there is no GPU: the product has no CPU fallback
    yield c
    c.close()
