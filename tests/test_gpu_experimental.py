"""GPU parity checks for the two code paths round 1 wrote after its GPU budget was spent.  Both ran green on a B200 in round 2
(profiles/r2a_first_run.log, r2b_second_run.log): the fused DCT-IV kernel became the default (12 % -> 63 % of its roofline);
the three-pass 2-D plan measured no faster than the default (1.249 ms against 1.247 ms at 8192 x 8192) and stays a knob.
Each case runs in its own process because the knobs are read once per process.

  SFC_DCT4_FUSED=1   TM_FAST_DCT4: DCT-IV / DST-IV rows in one kernel on the n/2-point complex transform (fft_tile.cuh)
  SFC_FFT2_TILE2D=1  TM_FAST_2D: three-pass plan for fft2 8192 x 8192 (tools/fft2_three_pass_emulation.py)
"""
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DCT4 = r'''
import numpy as np, sys
sys.path.insert(0, %r)
import scirs_b200 as sb
from oracle import consumers_oracle as co
rng = np.random.default_rng(4)
worst = 0.0
for n in (128, 256, 1024, 4096):
    for rows in (1, 64, 96):
        x = rng.standard_normal((rows, n))
        for norm in (None, "ortho"):
            for name, fn, ofn in (("dct", sb.dctn, co.dctn), ("idct", sb.idctn, co.idctn), ("dst", sb.dstn, co.dstn), ("idst", sb.idstn, co.idstn)):
                got = fn(x, 4, norm, [1])
                ref = ofn(x[: min(rows, 8)], 4, norm, [1])   # the oracle evaluates the O(n^2) sums literally
                g = got[: ref.shape[0]]
                e = np.linalg.norm(g - ref) / np.linalg.norm(ref)
                worst = max(worst, e)
                assert e <= 1e-12, (name, n, rows, norm, e)
# long rows (128 KiB tiles): two oracle rows, and the involution DCT-IV(DCT-IV(x)) = (n/2) x (orthonormal: x) on all of them
for n in (8192, 16384):
    x = rng.standard_normal((33, n))
    for name, fn, ofn in (("dct", sb.dctn, co.dctn), ("dst", sb.dstn, co.dstn)):
        got = fn(x, 4, "ortho", [1])
        ref = ofn(x[:2], 4, "ortho", [1])
        e = np.linalg.norm(got[:2] - ref) / np.linalg.norm(ref)
        back = fn(got, 4, "ortho", [1])
        e2 = np.linalg.norm(back - x) / np.linalg.norm(x)
        worst = max(worst, e, e2)
        assert e <= 3e-12 and e2 <= 3e-12, (name, n, e, e2)
# strided axes (column tiles) and all axes of a 3-D array
for shape, axes in (((256, 64), [0]), ((128, 1024, 8), [1]), ((128, 128, 128), None), ((512, 96), [0, 1])):
    x = rng.standard_normal(shape)
    for norm in (None, "ortho"):
        for name, fn, ofn in (("dct", sb.dctn, co.dctn), ("idst", sb.idstn, co.idstn)):
            got, ref = fn(x, 4, norm, axes), ofn(x, 4, norm, axes)
            e = np.linalg.norm(got - ref) / np.linalg.norm(ref)
            worst = max(worst, e)
            assert e <= 2e-12, (name, shape, axes, norm, e)
print("dct4 fused parity ok, worst rel-L2", worst)
''' % ROOT


def test_dct4_fused_kernel(build_artifacts):
    env = dict(os.environ, SFC_DCT4_FUSED="1")
    r = subprocess.run([sys.executable, "-c", DCT4], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "parity ok" in r.stdout


FFT2 = r'''
import numpy as np, sys, time
sys.path.insert(0, %r)
import scirs_b200 as sb
rng = np.random.default_rng(5)
n = 8192
x = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
p = sb.FftPlan([n, n], [0, 1])
d = p.describe()
assert "2-D three-pass" in d, d
got = p.execute(x).reshape(n, n)
ref = np.fft.fft2(x)   # an independent transform is enough for a first run; the oracle needs minutes at this size
e = np.linalg.norm(got - ref) / np.linalg.norm(ref)
print(d)
assert e <= 1e-12, e
pi = sb.FftPlan([n, n], [0, 1], "c2c", "f64", False, 1.0 / (n * n))
assert "2-D three-pass" in pi.describe()
back = pi.execute(got).reshape(n, n)
e2 = np.linalg.norm(back - x) / np.linalg.norm(x)
assert e2 <= 1e-12, e2
print("fft2 three-pass parity ok, rel-L2", e, "round trip", e2)
''' % ROOT


def test_fft2_three_pass_plan(build_artifacts):
    env = dict(os.environ, SFC_FFT2_TILE2D="1")
    r = subprocess.run([sys.executable, "-c", FFT2], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "parity ok" in r.stdout
