import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/liboracle.so), built on demand.  Test infrastructure only."""
    import oracle_lib
    return oracle_lib.load()


@pytest.fixture(scope="session")
def ctx():
    import torch  # noqa: F401  (device memory / streams plumbing; also verifies the GPU is visible)
    import mgf_b200
    c = mgf_b200.Context(device=0)
    yield c
    c.close()
