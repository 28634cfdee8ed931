import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_device_count() -> int:
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    if _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (GPU tests run under gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def capi():
    """The product library.  On a GPU box a missing/unloadable library is a hard failure."""
    from sdrainer_b200 import _build, capi as cap
    # both libraries are brought up to date BEFORE the first dlopen (no-ops when they are newer than their sources):
    # a rebuild of libsdrgpu.so after it has been loaded would put a second copy of it behind libsdrhost.so
    _build.build()
    _build.build_host()
    cap.lib()
    return cap


GOLDEN = os.path.join(ROOT, "tests", "golden")
