import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a CUDA device: gpu-marked tests are skipped, not failed."""
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items:
        return
    try:
        from scirs_b200 import _lib

        ndev = _lib.load().sfc_device_count()
    except Exception:
        ndev = 0
    if ndev == 0:
        skip = pytest.mark.skip(reason="no CUDA device")
        for it in gpu_items:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def build_artifacts():
    """Make sure the C-ABI library and the oracle engine exist (built by __graft_entry__.build())."""
    lib = os.path.join(ROOT, "scirs_b200", "lib", "libscirs2_fft_cuda.so")
    eng = os.path.join(ROOT, "oracle", "librustfft_port.so")
    if not (os.path.exists(lib) and os.path.exists(eng)):
        import __graft_entry__ as g

        g.build()
    return lib, eng
