import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


def pytest_sessionstart(session):
    """The built artefacts are git-ignored: in a fresh checkout build them before the first test instead of
    depending on the order in which the driver runs build() and pytest.  (On the GPU box they arrive with the
    snapshot, so nothing is compiled there.)"""
    import shutil
    lib_missing = not os.path.exists(os.path.join(ROOT, "metacache_b200", "libmcb200.so"))
    ref_here = os.path.exists("/root/reference/src/main.cpp")
    c1_missing = ref_here and not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "c1", "cli_cpu_reference.out"))
    if (lib_missing or c1_missing) and shutil.which("nvcc"):
        import __graft_entry__
        __graft_entry__.build()


def _has_gpu():
    try:
        from metacache_b200 import _lib
        return _lib.lib().mcb200_device_count() > 0
    except Exception:
        return False


HAS_GPU = None


def pytest_collection_modifyitems(config, items):
    global HAS_GPU
    if HAS_GPU is None:
        HAS_GPU = _has_gpu()
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
