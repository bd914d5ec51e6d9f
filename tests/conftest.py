import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def hosttest_lib():
    """ctypes handle on tests/native/libhosttest.so (kernel templates compiled for the host)."""
    import ctypes

    so = os.path.join(ROOT, "tests", "native", "libhosttest.so")
    src = [os.path.join(ROOT, "tests", "native", "host_ff.cpp")]
    deps = src + [os.path.join(ROOT, "snark_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "snark_b200", "csrc")) if f.endswith((".cuh", ".h"))]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so] + src)
    return ctypes.CDLL(so)
