import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ope():
    import ope_pkg
    return ope_pkg.load()


@pytest.fixture(scope="session")
def synth(ope):
    from ope_b200 import synth as s
    return s


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (checker only)."""
    import orc_py
    orc_py.lib()
    return orc_py


@pytest.fixture(scope="session")
def cuda_lib(ope):
    from ope_b200 import cuda_lib as c
    return c


@pytest.fixture(scope="session")
def ctx(cuda_lib):
    """A CUDA context through the C ABI. Fails loudly (no skip, no fallback) when the library or device is missing."""
    c = cuda_lib.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def model(synth):
    return synth.make_model(157825)


@pytest.fixture(scope="session")
def small_model(synth):
    return synth.make_model(20000)


def rot_trans_err(synth, A, B):
    return synth.pose_error(A, B)
