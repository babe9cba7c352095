import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _gpu_count():
    try:
        from transport_analysis_b200 import _lib

        return _lib.device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    if _gpu_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this environment")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_run.npz"))


@pytest.fixture(scope="session")
def emu():
    """CPU thread-emulation build of the kernels' per-CTA phase functions."""
    import ctypes

    src = os.path.join(ROOT, "tests", "emu", "emu.cpp")
    out_dir = os.path.join(ROOT, "tests", "emu", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libemu.so")
    deps = [src] + [os.path.join(ROOT, "transport_analysis_b200", "csrc", f)
                    for f in ("ta_common.cuh", "fft_plan.h", "fft_core.cuh", "windowed_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    return ctypes.CDLL(so)
