import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ORACLE_PATH = os.path.join(ROOT, "oracle", "libcannon_oracle.so")
CUDA_PATH = os.path.join(ROOT, "cannon_physics_b200", "libcannon_cuda.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _ensure_built():
    if not os.path.exists(ORACLE_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    if not os.path.exists(CUDA_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "cannon_physics_b200", "csrc")])


@pytest.fixture(scope="session")
def oracle_lib():
    """The CPU restatement (checker). Only tests / smoke / bench cpu_baseline may load it."""
    _ensure_built()
    from cannon_physics_b200 import _ffi
    return _ffi.bind(ORACLE_PATH)


@pytest.fixture(scope="session")
def cuda_lib():
    _ensure_built()
    import cannon_physics_b200 as cp
    assert cp.lib.cannon_backend() == b"cuda"
    return cp.lib
