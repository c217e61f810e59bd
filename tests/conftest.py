import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def build_artifacts():
    """Make sure the C-ABI library and the oracle engine exist (built by __graft_entry__.build())."""
    lib = os.path.join(ROOT, "scirs_b200", "lib", "libscirs2_fft_cuda.so")
    eng = os.path.join(ROOT, "oracle", "librustfft_port.so")
    if not (os.path.exists(lib) and os.path.exists(eng)):
        import __graft_entry__ as g

        g.build()
    return lib, eng
