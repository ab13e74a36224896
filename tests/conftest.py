import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """the CPU oracle (test infrastructure; see oracle/gat_oracle.h)"""
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def ctx():
    """one gat_b200 context on cuda:0 -- fails (not skips) when the library or the GPU is missing"""
    from gat_b200 import device
    c = device.Context(0)
    yield c
    c.close()
