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
    {"SFC_BLUE3_MIN": "32768", "SFC_THREE_LEVEL_MIN": "32768"},          # five-pass Bluestein, three-level rows
    {"SFC_GPIPE": "1", "SFC_GPIPE_MIN_TILES": "1"},                      # group-pipelined flavour on every eligible row pass
    {"SFC_PIPE_LATE": "2", "SFC_PIPE_LATE_MIN_TILES": "1"},              # late-prefetch persistent flavour on every eligible row pass
    {"SFC_ROW_LANE_GROUPS": "0"},                                        # 512 x 4 / 1024 x 2 row tiles with CTA-wide barriers
]


def test_late_prefetch_flavour(build_artifacts):
    """TM_PIPE_LATE: persistent CTAs, the next tile lands in the idle exchange buffer (cp.async.bulk + mbarrier) during the
    tail of the current one.  Every compiled shape, tile counts that are not multiples of the grid, both directions."""
    code = r'''
import numpy as np
from scirs_b200 import FftPlan
rng = np.random.default_rng(15)
def rel(a, b): return np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel())
for prec, n, rows, tol in (("f64", 4096, 1111, 1e-12), ("f64", 4096, 7, 1e-12), ("f64", 8192, 613, 1e-12), ("f64", 2048, 1402, 1e-12),
                           ("f64", 2048, 1401, 1e-12), ("f32", 8192, 901, 1e-5)):
    a = rng.standard_normal((rows, n)) + 1j * rng.standard_normal((rows, n))
    if prec == "f32":
        a = a.astype(np.complex64)
    for fwd in (True, False):
        p = FftPlan([rows, n], [1], "c2c", prec, fwd, 0.5)
        d = p.describe()
        assert "late-prefetch" in d, d
        got = p.execute(a).reshape(rows, n)
        ref = (np.fft.fft(a.astype(np.complex128), axis=1) if fwd else np.fft.ifft(a.astype(np.complex128), axis=1) * n) * 0.5
        e = rel(got.astype(np.complex128), ref)
        assert e < tol, (prec, n, rows, fwd, e)
        got2 = p.execute(a).reshape(rows, n)   # a second launch of the persistent kernel (fresh mbarrier phase)
        assert np.array_equal(got, got2)
print("late prefetch ok")
'''
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, SFC_PIPE_LATE="2", SFC_PIPE_LATE_MIN_TILES="1"),
                       capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0 and "late prefetch ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]



@pytest.mark.parametrize("env", KNOBS, ids=lambda e: ",".join(f"{k}={v}" for k, v in e.items()) or "defaults")
def test_knob_combinations_keep_parity(env, build_artifacts):
    r = subprocess.run([sys.executable, "-c", CODE], env=dict(os.environ, **env), capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_three_level_decomposition_small_and_full_size(build_artifacts):
    """n > lmax^2 = 2^26 goes three levels deep (plan.cu add_three_level).  Forced at small lengths against numpy, then
    one real 2^27-point transform checked by a round trip (1e-12) and on sampled bins against direct sums."""
    code = r'''
import numpy as np
from scirs_b200 import FftPlan
rng = np.random.default_rng(3)
for lg, b in ((15, 3), (17, 2), (20, 2), (22, 1)):
    n = 1 << lg
    a = rng.standard_normal((b, n)) + 1j * rng.standard_normal((b, n))
    p = FftPlan([b, n], [1])
    assert "three-level" in p.describe(), p.describe()
    got = p.execute(a).reshape(b, n)
    ref = np.fft.fft(a, axis=1)
    e = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    inv = FftPlan([b, n], [1], "c2c", "f64", False, 1.0 / n).execute(got).reshape(b, n)
    e2 = np.linalg.norm(inv - a) / np.linalg.norm(a)
    print(lg, e, e2)
    assert e < 1e-12 and e2 < 1e-12
'''
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, SFC_THREE_LEVEL_MIN="32768"), capture_output=True,
                       text=True, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    code2 = r'''
import numpy as np, torch
from scirs_b200 import FftPlan
from oracle import scirs2_fft_oracle as orc
n = 1 << 27
g = torch.Generator(device="cuda").manual_seed(5)
x = torch.randn(2 * n, dtype=torch.float64, device="cuda", generator=g)
y = torch.empty_like(x); z = torch.empty_like(x)
p = FftPlan([n], [0])
assert "three-level" in p.describe(), p.describe()
s = torch.cuda.current_stream().cuda_stream
p.execute_device(x, y, s)
FftPlan([n], [0], "c2c", "f64", False, 1.0 / n).execute_device(y, z, s)
torch.cuda.synchronize()
rt = float(torch.linalg.vector_norm(z - x) / torch.linalg.vector_norm(x))
xc = torch.view_as_complex(x.view(n, 2)).cpu().numpy()
bins = [1, 12345, n // 2 + 7, n - 1]
j = np.arange(n, dtype=np.int64)
ref = np.array([np.dot(xc, np.exp(-2j * np.pi * ((j * k) % n) / n)) for k in bins])   # exact integer phase reduction, f64 sum
got = torch.view_as_complex(y.view(n, 2))[bins].cpu().numpy()
e = np.abs(got - ref).max() / np.abs(ref).max()
print("round trip", rt, "sampled bins", e)
assert rt < 1e-12 and e < 1e-9
'''
    r = subprocess.run([sys.executable, "-c", code2], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
