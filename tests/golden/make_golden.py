"""Generate tests/golden/golden_v1.npz — seeded inputs and expected outputs for the hot path.

Two kinds of vectors:
  * kat_*  : the reference's own known-answer tests restated analytically
             (scirs2-fft/src/bin/accuracy_comparison.rs:83-267, rfft.rs:926-1032,
             planning.rs:733-754, fft/algorithms.rs doctests).  Expected values are closed
             forms — they do not depend on any FFT implementation.
  * ora_*  : seeded random inputs with outputs computed by an O(n^2) extended-precision DFT
             (oracle.dft_longdouble) composed per the reference's wrapper semantics.  These pin
             the oracle (and through it the GPU path) at non-power-of-two lengths, norm strings
             and axes subsets that the reference's own tests never touch.
The reference itself (Rust + un-vendored rustfft) cannot be run here, so there are no
reference-generated outputs; see DESIGN.md "Oracle".

Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import scirs2_fft_oracle as orc  # noqa: E402

out = {}
rng = np.random.default_rng(20261017)


def ld_fft(x, inverse=False):
    return orc.dft_longdouble(np.asarray(x, dtype=np.complex128), inverse)


def ld_fftn(x, axes, inverse=False):
    y = np.asarray(x, dtype=np.complex128)
    for a in axes:
        y = np.moveaxis(ld_fft(np.moveaxis(y, a, -1), inverse), -1, a)
    return y


# ---- analytic known answers -------------------------------------------------
# accuracy_comparison.rs:83-121: pure sine k=5 -> -i*N/2 at +k, +i*N/2 at -k
for n in (64, 128, 256, 512, 1024):
    t = np.arange(n)
    x = np.sin(2 * np.pi * 5 * t / n)
    exp = np.zeros(n, dtype=np.complex128)
    exp[5] = -0.5j * n
    exp[n - 5] = 0.5j * n
    out[f"kat_sine_{n}_in"] = x
    out[f"kat_sine_{n}_out"] = exp
# accuracy_comparison.rs:160-200: x[i] = sin(i) + i*cos(i/2) round trip (input is its own expected ifft(fft))
for n in (64, 256, 1024):
    i = np.arange(n)
    out[f"kat_roundtrip_{n}_in"] = np.sin(i) + 1j * np.cos(i / 2)
# accuracy_comparison.rs:202-267: 2-D sine at (3, 2)
for n in (16, 32, 64):
    yy, xx = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    a = np.sin(2 * np.pi * (3 * yy + 2 * xx) / n)
    exp = np.zeros((n, n), dtype=np.complex128)
    exp[3, 2] = -0.5j * n * n
    exp[n - 3, n - 2] = 0.5j * n * n
    out[f"kat_sine2d_{n}_in"] = a
    out[f"kat_sine2d_{n}_out"] = exp
# planning.rs:733-754: impulse -> flat unit spectrum
imp = np.zeros(8, dtype=np.complex128)
imp[0] = 1
out["kat_impulse_in"] = imp
out["kat_impulse_out"] = np.ones(8, dtype=np.complex128)
# rfft.rs:995-1032: sine N=16 k=2 -> |Im X[2]| = 8
out["kat_rsine16_in"] = np.sin(2 * np.pi * 2 * np.arange(16) / 16)
exp = np.zeros(9, dtype=np.complex128)
exp[2] = -8j
out["kat_rsine16_out"] = exp
# doctests: fft([1,2,3,4]) (algorithms.rs:117-130), fft2([[1,2],[3,4]]) (:280-292)
out["kat_1234_in"] = np.array([1.0, 2.0, 3.0, 4.0])
out["kat_1234_out"] = np.array([10, -2 + 2j, -2, -2 - 2j], dtype=np.complex128)
out["kat_2x2_in"] = np.array([[1.0, 2.0], [3.0, 4.0]])
out["kat_2x2_out"] = np.array([[10, -2], [-4, 0]], dtype=np.complex128)

# ---- extended-precision vectors under the reference's wrapper semantics -------
for n in (3, 5, 7, 12, 17, 100, 127, 243, 1000):
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    out[f"ora_fft_n{n}_in"] = x
    out[f"ora_fft_n{n}_out"] = ld_fft(x)  # fft(x, Some(n))
    out[f"ora_ifft_n{n}_out"] = ld_fft(x, True) / n
# fft(x, None): pads to the next power of two (algorithms.rs:142)
x = rng.standard_normal(100)
out["ora_fft_pad_in"] = x
out["ora_fft_pad_out"] = ld_fft(np.concatenate([x, np.zeros(28)]))
# ifft(x, None) on non-pow2 length: size 128, scale 1/128, truncated to 100 (algorithms.rs:221,258-260)
xc = rng.standard_normal(100) + 1j * rng.standard_normal(100)
out["ora_ifft_pad_in"] = xc
out["ora_ifft_pad_out"] = (ld_fft(np.concatenate([xc, np.zeros(28)]), True) / 128)[:100]
# rfft (no pow2 padding, rfft.rs:45) and irfft
x = rng.standard_normal(90)
out["ora_rfft_in"] = x
out["ora_rfft_out"] = ld_fft(x)[:46]
out["ora_irfft_out"] = x  # irfft(rfft(x), Some(90)) == x
# fftn on 6x10x12 with axes subsets and every norm string: forward scale uses ALL dims (algorithms.rs:694)
v = rng.standard_normal((6, 10, 12))
out["ora_fftn_in"] = v
total = v.size
for tag, axes in (("all", [0, 1, 2]), ("a20", [2, 0]), ("a1", [1])):
    base = ld_fftn(v, axes)
    out[f"ora_fftn_{tag}_none"] = base
    out[f"ora_fftn_{tag}_backward"] = base / total
    out[f"ora_fftn_{tag}_ortho"] = base / np.sqrt(total)
    out[f"ora_fftn_{tag}_forward"] = base / total
    ib = ld_fftn(v, axes, True)
    sub = np.prod([v.shape[a] for a in axes])  # ifftn: listed axes only (algorithms.rs:876)
    out[f"ora_ifftn_{tag}_backward"] = ib / sub
    out[f"ora_ifftn_{tag}_ortho"] = ib / np.sqrt(sub)
    out[f"ora_ifftn_{tag}_forward"] = ib
# fft2 with padding/cropping shape and norm (algorithms.rs:334-347,385-395)
a = rng.standard_normal((9, 14))
out["ora_fft2_in"] = a
pad = np.zeros((12, 10))
pad[:9, :10] = a[:9, :10]
out["ora_fft2_shape_12x10"] = ld_fftn(pad, [1, 0])
out["ora_fft2_ortho"] = ld_fftn(a, [1, 0]) / np.sqrt(a.size)
out["ora_ifft2_default"] = ld_fftn(a, [1, 0], True) / a.size

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
np.savez_compressed(path, **out)
print("wrote", path, len(out), "arrays", os.path.getsize(path), "bytes")
