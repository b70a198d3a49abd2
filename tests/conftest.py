import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle_lib import load_oracle
    return load_oracle()


@pytest.fixture(scope="session")
def cuda():
    """The product library; GPU tests fail loudly if it is missing or no device is usable."""
    import sgdnet_b200
    lib = sgdnet_b200.product()
    import ctypes
    cnt = ctypes.c_int(0)
    rc = lib.sym("device_count")(ctypes.byref(cnt))
    assert rc == 0 and cnt.value > 0, "no CUDA device: the GPU tests need a B200"
    return lib


def golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
