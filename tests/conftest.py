import os
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parents[1]
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

GOLDEN = REPO / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200 box)")


def _has_gpu():
    try:
        import ctypes
        cuda = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        return cuda.cuInit(0) == 0 and cuda.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


# AKUA_TESTS_ON_EMULATOR=1 (diagnostic, with AKUA_PBF_LIB=tests/emu/_build/libakua_pbf_emu.so): run the logic of the small -m gpu
# tests against the host solver compiled for the SIMT emulator — catches host-side regressions without a GPU; the verdict
# that counts is still the run on a B200.
HAS_GPU = _has_gpu() or os.environ.get("AKUA_TESTS_ON_EMULATOR") == "1"


def pytest_collection_modifyitems(config, items):
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this environment")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def akua_lib():
    from akuaengine_b200 import build_library, load_library
    build_library()
    return load_library()
