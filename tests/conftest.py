import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def offsets():
    return np.load(os.path.join(GOLDEN, "brief_offsets.npy")).astype(np.int32)


@pytest.fixture(scope="session")
def kitti():
    import cv2
    img = cv2.imread(os.path.join(GOLDEN, "kitti_frame.png"), 0)
    assert img is not None and img.shape == (376, 1241)
    return img


@pytest.fixture(scope="session")
def emul():
    """Host build of the device arithmetic (tests/emul/emul_core.cpp)."""
    import ctypes as C
    d = os.path.join(ROOT, "tests", "emul")
    so = os.path.join(d, "libyavo_emul.so")
    srcs = [os.path.join(d, "emul_core.cpp"), os.path.join(ROOT, "ya_vo_b200", "csrc", "fast_core.h"),
            os.path.join(ROOT, "ya_vo_b200", "csrc", "select_serial.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17", "-o", so, srcs[0]])
    L = C.CDLL(so)
    L.emul_harris.restype = C.c_float
    L.emul_score_from_tensor.restype = C.c_float
    return L


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library; GPU tests fail loudly if it is missing or no device is present."""
    from ya_vo_b200 import capi
    capi.build()
    return capi
