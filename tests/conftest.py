import os as _os

# slab ranks emulated in ONE process (SlabCluster(peer=True)) spin on device flags: every kernel must be loaded before anyone waits, so that
# a first launch by one rank never has to load a module while another rank's wait kernel is resident (read at CUDA initialisation)
_os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, 'tests')):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_device_present():
    """True when the C-ABI library can create a simulator, i.e. a CUDA device is visible (no torch import needed)."""
    try:
        import ctypes
        for name in ("libcuda.so.1", "libcuda.so"):
            try:
                cu = ctypes.CDLL(name)
                break
            except OSError:
                cu = None
        if cu is None or cu.cuInit(0) != 0:
            return False
        n = ctypes.c_int(0)
        return cu.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    except Exception:                       # noqa: BLE001
        return False


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped, not failed, on a machine without a CUDA device (the library itself has no CPU path)."""
    if _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device: libsoftmac_b200 has no CPU path")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
