"""GPU: the planner's experiment knobs keep producing the oracle's numbers — every selectable code
path stays parity-checked even when it is off by default."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CODE = r"""
import numpy as np, scirs_b200 as sb
from scirs_b200 import FftPlan
rng = np.random.default_rng(5)
def c(*s): return rng.standard_normal(s) + 1j * rng.standard_normal(s)
def rel(a, b): return np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel())
worst = 0.0
for shape, axes in (([600, 4096], [1]), ([320, 8192], [1]), ([4, 1 << 16], [1]), ([24, 20011], [1]), ([40, 512, 16], [1]),
                    ([16, 256, 256], [1, 2]), ([2400, 1024], [1])):
    a = c(*shape)
    got = FftPlan(shape, axes).execute(a).reshape(shape)
    worst = max(worst, rel(got, np.fft.fftn(a, axes=axes)))
    inv = FftPlan(shape, axes, "c2c", "f64", False).execute(a).reshape(shape)
    worst = max(worst, rel(inv, np.fft.ifftn(a, axes=axes, norm="forward")))
print("worst", worst)
assert worst < 1e-12, worst
"""

KNOBS = [
    {},
    {"SFC_PIPE": "1", "SFC_PIPE_MIN_TILES": "1"},                      # TMA-pipelined flavour on every eligible pass
    {"SFC_PIPE": "2", "SFC_PIPE_MIN_TILES": "1", "SFC_PIPE_BIG": "0"},  # rows only, big-row pipelining off
    {"SFC_L2_CHUNK_MB": "1", "SFC_L2_WAYS": "3", "SFC_L2_TOTAL_MB": "8"},  # L2-blocked rounds on side streams
    {"SFC_ROW_FOURSTEP": "4096", "SFC_L2_CHUNK_MB": "2", "SFC_L2_WAYS": "2", "SFC_L2_TOTAL_MB": "8"},
    {"SFC_CHIRP_GEN": "0", "SFC_TILE_GROUP_LOG2": "0"},                  # table-driven Bluestein, plain CTA order
    {"SFC_BLUE_L1": "64", "SFC_TILE_GROUP_LOG2": "5"},
    {"SFC_FAST": "0"},                                                   # generic flavour everywhere
    {"SFC_FORCE_E": "8"},
    {"SFC_WORK_MB": "1"},                                                # many rounds through a tiny work area
]


@pytest.mark.parametrize("env", KNOBS, ids=lambda e: ",".join(f"{k}={v}" for k, v in e.items()) or "defaults")
def test_knob_combinations_keep_parity(env, build_artifacts):
    r = subprocess.run([sys.executable, "-c", CODE], env=dict(os.environ, **env), capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
