"""Generate tests/golden/golden_consumers_v1.npz — seeded inputs and EXTENDED-PRECISION expected outputs for the
consumers of the hot path: the reference's DCT/DST sums (dct.rs:425-757, dst.rs:409-702) evaluated
in numpy longdouble (eps 1.1e-19) with exactly reduced angles, and the Hartley / hfft / ihfft / hilbert compositions
(hartley.rs:37-66, hfft/*.rs, lib.rs:437-516) on top of the extended-precision DFT of the core oracle.  They pin
oracle/consumers_oracle.py (CPU test) and, through the same file, the GPU product (GPU test) independently of f64
rounding in either.  The reference itself cannot be run here (no Rust toolchain): these are not reference outputs.

Run:  python tests/golden/make_golden_consumers.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import scirs2_fft_oracle as orc  # noqa: E402

LD = np.longdouble
PI = LD("3.14159265358979323846264338327950288419716939937510")
rng = np.random.default_rng(20261018)
out = {}


def trig(num, den, f):
    """f(pi * num / den) with the integer numerator reduced mod 2*den first (exact)."""
    num = np.asarray(num, dtype=np.int64) % (2 * den)
    ang = PI * num.astype(LD) / LD(den)
    return np.cos(ang) if f == "cos" else np.sin(ang)


def sums(x, f, a2, b2, D):
    """sum_i x[i] f(pi (2i + a2)(2k + b2) / (4D)) for k < n, in long double."""
    n = x.size
    i = np.arange(n, dtype=np.int64)[None, :]
    k = np.arange(n, dtype=np.int64)[:, None]
    return (trig((2 * i + a2) * (2 * k + b2), 4 * D, f) * x.astype(LD)[None, :]).sum(axis=1)


def sq(v):
    return np.sqrt(LD(v))


def dct_ref(x, t, inverse, ortho):
    """The reference's DCT family, term by term (dct.rs), in long double."""
    n = x.size
    X = x.astype(LD).copy()
    if t == 1:
        m = n - 1
        if not inverse:
            r = sums(X, "cos", 0, 0, m)
            r[0] *= LD(0.5); r[-1] *= LD(0.5)
            if ortho:
                r *= sq(2) / sq(m); r[0] /= sq(2); r[-1] /= sq(2)
            return r
        if ortho:
            X *= sq(m) / sq(2); X[0] *= sq(2); X[-1] *= sq(2)
        X[0] *= LD(0.5); X[-1] *= LD(0.5)
        return sums(X, "cos", 0, 0, m) * LD(2) / LD(m)
    if t == 2:
        if not inverse:
            r = sums(X, "cos", 1, 0, n)
            if ortho:
                r *= sq(2) / sq(n); r[0] /= sq(2)
            return r
        if ortho:
            X *= sq(n) / sq(2); X[0] *= sq(2)
        X[0] *= LD(0.5)
        return sums(X, "cos", 0, 1, n) * LD(2) / LD(n)
    if t == 3:
        if not inverse:
            if ortho:
                X *= sq(n) / sq(2); X[0] /= sq(2)
            X[0] *= LD(0.5)
            return sums(X, "cos", 0, 1, n) * LD(2) / LD(n)
        if ortho:
            X *= sq(2) / sq(n); X[0] *= sq(2)
        return sums(X, "cos", 0, 1, n)
    if not inverse:
        r = sums(X, "cos", 1, 1, n)
        return r * sq(2) / sq(n) if ortho else r
    X *= (sq(n) / sq(2)) if ortho else LD(2) / LD(n)
    r = sums(X, "cos", 1, 1, n)
    return r * sq(2) / sq(n) if ortho else r


def dst_ref(x, t, inverse, ortho):
    n = x.size
    X = x.astype(LD).copy()
    if t == 1:
        m = n + 1
        if not inverse:
            return sums(X, "sin", 2, 2, m) * ((sq(2) / sq(m)) if ortho else LD(2) / sq(m))
        return sums(X * sq(m) / LD(2), "sin", 2, 2, m) * LD(2) / sq(m)
    if t == 2:
        if not inverse:
            r = sums(X, "sin", 1, 2, n)
            return r * sq(2) / sq(n) if ortho else r
        if ortho:
            X *= sq(n) / sq(2)
        return sums(X, "sin", 2, 1, n) * LD(0.5)
    if t == 3:
        if not inverse:
            return sums(X, "sin", 2, 1, n) * ((sq(2) / sq(n) / LD(2)) if ortho else LD(0.5))
        X *= (sq(n) / sq(2) * LD(2)) if ortho else LD(2)
        return sums(X, "sin", 1, 2, n)
    if not inverse:
        return sums(X, "sin", 1, 1, n) * ((sq(2) / sq(n)) if ortho else LD(2))
    X *= (sq(n) / sq(2)) if ortho else LD(0.5)
    return sums(X, "sin", 1, 1, n) * LD(2)


for n in (2, 5, 8, 16, 33, 128, 257):
    x = rng.standard_normal(n)
    out[f"trig_x_{n}"] = x
    for kind, fn in (("dct", dct_ref), ("dst", dst_ref)):
        for t in (1, 2, 3, 4):
            for inv in (0, 1):
                for ortho in (0, 1):
                    out[f"{kind}_{n}_t{t}_i{inv}_o{ortho}"] = fn(x, t, bool(inv), bool(ortho)).astype(np.float64)

for n in (1, 4, 5, 12, 64, 100):
    x = rng.standard_normal(n)
    z = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    out[f"real_{n}"] = x
    out[f"cplx_{n}"] = z
    P = orc.next_power_of_two(n)
    xp = np.zeros(P); xp[:n] = x
    F = orc.dft_longdouble(xp.astype(np.complex128))
    out[f"dht_{n}"] = (F[:n].real - F[:n].imag).astype(np.float64)                      # hartley.rs:57-62
    zz = z.copy(); zz[0] = zz[0].real
    m = n + 3
    zp = np.zeros(m, dtype=np.complex128); zp[:n] = zz
    out[f"hfft_{n}_n{m}"] = orc.dft_longdouble(zp).real.astype(np.float64)              # complex_to_real.rs:113-135
    xp2 = np.zeros(m); xp2[:n] = x
    r = orc.dft_longdouble(xp2.astype(np.complex128), inverse=True) / m
    ih = np.empty(m, dtype=np.complex128)
    mid = (m + 1) // 2
    ih[0] = r[0].real; ih[1:mid] = r[1:mid]; ih[mid:] = np.conj(r[m - mid:0:-1])        # real_to_complex.rs:128-147
    out[f"ihfft_{n}_n{m}"] = ih
    h = np.zeros(n, dtype=np.complex128); h[0] = 1
    if n % 2 == 0:
        h[n // 2] = 1; h[1:n // 2] = -2j
    else:
        h[1:(n + 1) // 2] = -2j
    S = F[:n] * h                                                                        # lib.rs:470-510
    Sp = np.zeros(P, dtype=np.complex128); Sp[:n] = S
    out[f"hilbert_{n}"] = (orc.dft_longdouble(Sp, inverse=True) / P)[:n].astype(np.complex128)

np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_consumers_v1.npz"), **out)
print("wrote", len(out), "arrays")
