"""Developer parity sweep on a real GPU: every kernel instantiation + planner path vs the oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scirs_b200 as sb
from scirs_b200 import FftPlan
from oracle import scirs2_fft_oracle as orc

rng = np.random.default_rng(0)
bad = 0
def chk(name, got, ref, tol):
    global bad
    e = orc.rel_l2(got, ref)
    flag = "" if e <= tol else "   <-- FAIL"
    if e > tol: bad += 1
    print(f"{name:60s} rel_l2={e:.3e}{flag}", flush=True)

def cplx(*s):
    return rng.standard_normal(s) + 1j * rng.standard_normal(s)

import scipy.fft as sf
# 1. batched 1-D c2c rows, all pow2 lengths (ROW tiles), f64 + f32
for prec, tol, cd in (("f64", 1e-13, np.complex128), ("f32", 2e-6, np.complex64)):
    for lg in range(1, 15 if prec == "f32" else 14):
        n = 1 << lg
        b = max(3, min(300, (1 << 16) // n))
        x = cplx(b, n).astype(cd)
        p = FftPlan([b, n], [1], "c2c", prec, True)
        y = p.execute(x).reshape(b, n)
        chk(f"c2c rows {prec} n={n} b={b} fwd", y, sf.fft(x.astype(np.complex128), axis=1), tol)
        p = FftPlan([b, n], [1], "c2c", prec, False, 1.0 / n)
        y = p.execute(x).reshape(b, n)
        chk(f"c2c rows {prec} n={n} b={b} inv", y, sf.ifft(x.astype(np.complex128), axis=1), tol)
# 2. strided axis (COL tiles)
for prec, tol, cd in (("f64", 1e-13, np.complex128), ("f32", 2e-6, np.complex64)):
    for n in (2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192):
        inner = 24 if n <= 1024 else 5
        x = cplx(3, n, inner).astype(cd)
        p = FftPlan([3, n, inner], [1], "c2c", prec, True)
        y = p.execute(x).reshape(3, n, inner)
        chk(f"c2c cols {prec} n={n} inner={inner}", y, sf.fft(x.astype(np.complex128), axis=1), tol)
# 3. four-step
for n in (1 << 14, 1 << 15, 1 << 17, 1 << 20):
    x = cplx(3, n)
    y = FftPlan([3, n], [1], "c2c", "f64", True).execute(x).reshape(3, n)
    chk(f"four-step rows n={n}", y, sf.fft(x, axis=1), 1e-13)
    y = FftPlan([3, n], [1], "c2c", "f64", False, 1.0/n).execute(x).reshape(3, n)
    chk(f"four-step rows n={n} inv", y, sf.ifft(x, axis=1), 1e-13)
x = cplx(2, 1 << 14, 3)
y = FftPlan([2, 1 << 14, 3], [1], "c2c", "f64", True).execute(x).reshape(x.shape)
chk("four-step cols n=16384 inner=3", y, sf.fft(x, axis=1), 1e-13)
x = cplx(2, 1 << 15).astype(np.complex64)
y = FftPlan([2, 1 << 15], [1], "c2c", "f32", True).execute(x).reshape(x.shape)
chk("four-step rows f32 n=32768", y, sf.fft(x.astype(np.complex128), axis=1), 2e-6)
# 4. Bluestein
for n in (3, 5, 7, 12, 100, 1000, 4095, 4097, 6561, 10007, 100003):
    x = cplx(4, n)
    y = FftPlan([4, n], [1], "c2c", "f64", True).execute(x).reshape(4, n)
    chk(f"bluestein rows n={n}", y, sf.fft(x, axis=1), 1e-12)
    y = FftPlan([4, n], [1], "c2c", "f64", False, 1.0/n).execute(x).reshape(4, n)
    chk(f"bluestein rows n={n} inv", y, sf.ifft(x, axis=1), 1e-12)
x = cplx(3, 100, 7)
y = FftPlan([3, 100, 7], [1], "c2c", "f64", True).execute(x).reshape(x.shape)
chk("bluestein cols n=100 inner=7", y, sf.fft(x, axis=1), 1e-12)
x = cplx(2, 5000, 3)
y = FftPlan([2, 5000, 3], [1], "c2c", "f64", True).execute(x).reshape(x.shape)
chk("bluestein 3-pass cols n=5000 inner=3", y, sf.fft(x, axis=1), 1e-12)
# 5. real transforms
for prec, tol, rd in (("f64", 1e-13, np.float64), ("f32", 2e-6, np.float32)):
    for n in (64, 128, 256, 1024, 4096, 8192, 16384):
        x = rng.standard_normal((5, n)).astype(rd)
        y = FftPlan([5, n], [1], "r2c", prec).execute(x).reshape(5, n // 2 + 1)
        ref = sf.rfft(x.astype(np.float64), axis=1)
        chk(f"r2c fast {prec} n={n}", y, ref, tol)
        z = FftPlan([5, n], [1], "c2r", prec, scale=1.0 / n).execute(ref.astype(y.dtype)).reshape(5, n)
        chk(f"c2r fast {prec} n={n}", z, x.astype(np.float64), tol)
for n in (2, 6, 10, 30, 100, 1000, 32768):
    x = rng.standard_normal((3, n))
    y = FftPlan([3, n], [1], "r2c", "f64").execute(x).reshape(3, n // 2 + 1)
    chk(f"r2c general n={n}", y, sf.rfft(x, axis=1), 1e-12)
    z = FftPlan([3, n], [1], "c2r", "f64", scale=1.0 / n).execute(sf.rfft(x, axis=1)).reshape(3, n)
    chk(f"c2r general n={n}", z, x, 1e-12)
# 6. N-D
x = cplx(16, 32, 64)
chk("fftn 16x32x64", FftPlan(x.shape, None, "c2c").execute(x).reshape(x.shape), sf.fftn(x), 1e-13)
x = cplx(12, 10, 18)
chk("fftn 12x10x18 (bluestein axes)", FftPlan(x.shape, None, "c2c").execute(x).reshape(x.shape), sf.fftn(x), 1e-12)
x = rng.standard_normal((8, 16, 64))
chk("rfftn 8x16x64", FftPlan(x.shape, None, "r2c").execute(x).reshape(8, 16, 33), sf.rfftn(x), 1e-13)
chk("irfftn 8x16x64", FftPlan(x.shape, None, "c2r", scale=1.0/x.size).execute(sf.rfftn(x)).reshape(x.shape), x, 1e-13)
x = rng.standard_normal((6, 10, 14))
chk("rfftn 6x10x14", FftPlan(x.shape, None, "r2c").execute(x).reshape(6, 10, 8), sf.rfftn(x), 1e-12)
chk("irfftn 6x10x14", FftPlan(x.shape, None, "c2r", scale=1.0/x.size).execute(sf.rfftn(x)).reshape(x.shape), x, 1e-12)
# 7. drop-in API
x = rng.standard_normal(1000)
chk("api fft(real, None) pads to 1024", sb.fft(x), orc.fft(x), 1e-13)
chk("api ifft(None) truncates", sb.ifft(cplx(1000)[:]), orc.ifft(rng.standard_normal(1000)*0 + cplx(1000)), 10)  # shape check only
xc = cplx(1000)
chk("api ifft", sb.ifft(xc), orc.ifft(xc), 1e-13)
chk("api rfft n=None", sb.rfft(x), orc.rfft(x), 1e-12)
chk("api rfft n=2048", sb.rfft(x, 2048), orc.rfft(x, 2048), 1e-13)
s = orc.rfft(x[:512])
chk("api irfft fast", sb.irfft(s, 512), orc.irfft(s, 512), 1e-13)
chk("api irfft n=None", sb.irfft(s), orc.irfft(s), 1e-13)
chk("api irfft odd", sb.irfft(s, 501), orc.irfft(s, 501), 1e-12)
chk("api irfft short", sb.irfft(s, 100), orc.irfft(s, 100), 1e-12)
chk("api irfft long", sb.irfft(s, 2000), orc.irfft(s, 2000), 1e-12)
a = rng.standard_normal((20, 36))
for norm in (None, "backward", "ortho", "forward", "bogus"):
    chk(f"api fft2 norm={norm}", sb.fft2(a, None, None, norm), orc.fft2(a, None, None, norm), 1e-12)
    chk(f"api ifft2 norm={norm}", sb.ifft2(a, None, None, norm), orc.ifft2(a, None, None, norm), 1e-12)
chk("api fft2 shape pad", sb.fft2(a, (32, 32)), orc.fft2(a, (32, 32)), 1e-13)
chk("api fft2 shape crop", sb.fft2(a, (16, 40)), orc.fft2(a, (16, 40)), 1e-12)
chk("api rfft2", sb.rfft2(a), orc.rfft2(a), 1e-12)
chk("api irfft2", sb.irfft2(orc.rfft2(a)), orc.irfft2(orc.rfft2(a)), 1e-12)
v = rng.standard_normal((6, 8, 10))
for norm in (None, "backward", "ortho", "forward"):
    chk(f"api fftn axes=[2,0] norm={norm}", sb.fftn(v, None, [2, 0], norm), orc.fftn(v, None, [2, 0], norm), 1e-12)
    chk(f"api ifftn axes=[1] norm={norm}", sb.ifftn(v, None, [1], norm), orc.ifftn(v, None, [1], norm), 1e-12)
chk("api fftn dup axes", sb.fftn(v, None, [1, 1]), orc.fftn(v, None, [1, 1]), 1e-12)
chk("api fftn shape", sb.fftn(v, [8, 8, 8]), orc.fftn(v, [8, 8, 8]), 1e-12)
chk("api rfftn", sb.rfftn(v), orc.rfftn(v), 1e-12)
chk("api rfftn axes=[0,1]", sb.rfftn(v, None, [0, 1]), orc.rfftn(v, None, [0, 1]), 1e-12)
chk("api rfftn shape given", sb.rfftn(v, [6, 8, 16]), orc.rfftn(v, [6, 8, 16]), 1e-12)
sp = orc.rfftn(v)
chk("api irfftn", sb.irfftn(sp), orc.irfftn(sp), 1e-12)
chk("api irfftn shape", sb.irfftn(sp, [6, 8, 10]), orc.irfftn(sp, [6, 8, 10]), 1e-12)
chk("api irfftn axes=[2]", sb.irfftn(sp, None, [2]), orc.irfftn(sp, None, [2]), 1e-12)
chk("api irfftn odd pad", sb.irfftn(sp, [7, 9, 11]), orc.irfftn(sp, [7, 9, 11]), 1e-12)
chk("api fft_strided", sb.fft_strided(v, 1), orc.fft_strided(v, 1), 1e-12)
chk("api ifft_strided", sb.ifft_strided(v + 0j, 0), orc.ifft_strided(v + 0j, 0), 1e-12)
chk("api fft f32 input", sb.fft(x.astype(np.float32)), orc.fft(x.astype(np.float32)), 1e-13)
chk("api fft c64 input", sb.fft(xc.astype(np.complex64), 777), orc.fft(xc.astype(np.complex64), 777), 1e-12)
print("cache:", sb.get_global_cache().get_stats())
print("FAILURES:", bad)
sys.exit(1 if bad else 0)
