"""Chirp z-transform — mirror of scirs2-fft/src/czt.rs:24-360.

The reference file sets the whole Bluestein-style algorithm up and then stubs its three FFT calls out with zero
vectors (czt.rs:110-113, 239-252: "TODO: Fix FFT reference"), so its `czt` returns zeros.  This mirror keeps the
reference's names, arguments, defaults and error texts and runs the algorithm that file describes, on the GPU
(`sfc_czt`: two plan executions with every chirp multiply fused into them).  `czt` with default parameters equals
`fft` — the property the reference's own `test_czt_as_fft` (czt.rs:396-410) states.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from .error import ValueError_, check
from .fft import _ptr


def czt_points(m: int, a: Optional[complex] = None, w: Optional[complex] = None) -> np.ndarray:
    """czt.rs:24-40: the points of the z-plane the transform is evaluated at."""
    a = 1.0 + 0.0j if a is None else complex(a)
    k = np.arange(m, dtype=np.float64)
    if w is not None:
        return a * np.power(complex(w), -k)
    return a * np.exp(2j * np.pi * k / m)


class CZT:
    """czt.rs:45-275"""

    def __init__(self, n: int, m: Optional[int] = None, w: Optional[complex] = None, a: Optional[complex] = None):
        if n < 1:
            raise ValueError_("n must be positive")
        m = n if m is None else int(m)
        if m < 1:
            raise ValueError_("m must be positive")
        self.n, self.m = int(n), m
        self.w = None if w is None else complex(w)
        self.a = 1.0 + 0.0j if a is None else complex(a)

    def points(self) -> np.ndarray:
        return czt_points(self.m, self.a, self.w)

    def transform(self, x, axis: Optional[int] = None) -> np.ndarray:
        """czt.rs:143-224 (1-D and 2-D arrays, as there)."""
        arr = np.asarray(x, dtype=np.complex128)
        nd = arr.ndim
        if nd == 0:
            raise ValueError_("Invalid axis")
        ax = nd - 1 if axis is None else (nd + axis if axis < 0 else axis)
        if ax < 0 or ax >= nd:
            raise ValueError_("Invalid axis")
        if arr.shape[ax] != self.n:
            raise ValueError_(f"Input size ({arr.shape[ax]}) doesn't match CZT size ({self.n})")
        if nd > 2:
            raise ValueError_("CZT currently only supports 1D and 2D arrays")
        rows = np.ascontiguousarray(np.moveaxis(arr, ax, -1).reshape(-1, self.n))
        out = np.empty((rows.shape[0], self.m), dtype=np.complex128)
        lib = _lib.load()
        w = self.w if self.w is not None else 0j
        check(lib.sfc_czt(_ptr(rows), rows.shape[0], self.n, self.m, int(self.w is not None), w.real, w.imag, self.a.real,
                          self.a.imag, _ptr(out)))
        shape = list(np.moveaxis(arr, ax, -1).shape)
        shape[-1] = self.m
        return np.moveaxis(out.reshape(shape), -1, ax)


def czt(x, m: Optional[int] = None, w: Optional[complex] = None, a: Optional[complex] = None, axis: Optional[int] = None):
    """czt.rs:279-303"""
    arr = np.asarray(x)
    ax = arr.ndim - 1 if axis is None else (arr.ndim + axis if axis < 0 else axis)
    return CZT(arr.shape[ax], m, w, a).transform(arr, axis)


def zoom_fft(x, m: int, f0: float, f1: float, oversampling: Optional[float] = None):
    """czt.rs:315-360: m points of the spectrum between the normalised frequencies f0 < f1 in [0, 1]."""
    if not (0.0 <= f0 <= 1.0) or not (0.0 <= f1 <= 1.0):
        raise ValueError_("Frequencies must be in range [0, 1]")
    if f0 >= f1:
        raise ValueError_("f0 must be less than f1")
    oversampling = 2.0 if oversampling is None else float(oversampling)
    if oversampling < 1.0:
        raise ValueError_("Oversampling must be >= 1")
    arr = np.asarray(x)
    n = arr.shape[-1]
    k0 = f0 * n * oversampling
    k1 = f1 * n * oversampling
    step = (k1 - k0) / (m - 1)
    phi = 2.0 * np.pi * k0 / (n * oversampling)
    theta = -2.0 * np.pi * step / (n * oversampling)
    return czt(arr, m, complex(np.cos(theta), np.sin(theta)), complex(np.cos(phi), np.sin(phi)), arr.ndim - 1)
