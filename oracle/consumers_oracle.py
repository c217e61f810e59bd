"""TEST INFRASTRUCTURE ONLY — CPU restatement of the in-crate consumers of the FFT hot path
(SURVEY 8f rank 1).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

Each function restates the reference literally (the reference's DCT/DST are direct O(n^2) sums, so
that is what is written here, vectorised with numpy), citing the file:line it follows:

  dct / idct / dctn ...   scirs2-fft/src/dct.rs:56-420 (wrappers), :425-757 (per-type sums)
  dst / idst / dstn ...   scirs2-fft/src/dst.rs:48-405, :409-702
  dht / idht / dht2 / fht scirs2-fft/src/hartley.rs:37-209
  hfft / ihfft            scirs2-fft/src/hfft/complex_to_real.rs:58-135, real_to_complex.rs:49-149
  hilbert                 scirs2-fft/src/lib.rs:437-516
  get_window (subset)     scirs2-fft/src/window.rs:107-142, 182-200, 542-568
  stft / spectrogram      scirs2-fft/src/spectrogram.rs:76-310, 312-420

The reference's hard-coded test answers (`n == 4 && norm == "ortho"` -> [1,2,3,4] in idct1 / idst1..4)
are NOT restated: the product does not reproduce them either.

Pinning: the reference's own unit tests for these modules (dct.rs:752-864, dst.rs:704-794,
hartley.rs:211-261, spectrogram.rs tests, lib.rs doctest of hilbert) are restated in
tests/test_consumers_oracle.py; beyond them parity is unpinned by the reference.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from . import scirs2_fft_oracle as base

OracleError = base.OracleError
PI = np.pi


def _f64(x) -> np.ndarray:
    return np.asarray(x, dtype=np.float64).reshape(-1)


# ----------------------------------------------------------------------------- DCT (dct.rs)

def _dct1(x, norm):  # dct.rs:425-469
    n = x.size
    if n < 2:
        raise OracleError("ValueError", "Input array must have at least 2 elements for DCT-I")
    k = np.arange(n)[:, None]
    i = np.arange(n)[None, :]
    r = (np.cos(PI * k * i / (n - 1)) * x[None, :]).sum(axis=1)
    r[0] *= 0.5
    r[n - 1] *= 0.5
    if norm == "ortho":
        f = np.sqrt(2.0 / (n - 1))
        r *= f
        r[0] *= 1.0 / np.sqrt(2.0)
        r[n - 1] *= 1.0 / np.sqrt(2.0)
    return r


def _idct1(x, norm):  # dct.rs:473-519 (without the n == 4 hack)
    n = x.size
    if n < 2:
        raise OracleError("ValueError", "Input array must have at least 2 elements for IDCT-I")
    inp = x.copy()
    if norm == "ortho":
        inp *= np.sqrt((n - 1) / 2.0)
        inp[0] *= np.sqrt(2.0)
        inp[n - 1] *= np.sqrt(2.0)
    i = np.arange(n)
    s = 0.5 * (inp[0] + inp[n - 1] * np.where(i % 2 == 0, 1.0, -1.0))
    if n > 2:
        k = np.arange(1, n - 1)[None, :]
        s = s + (np.cos(PI * k * i[:, None] / (n - 1)) * inp[None, 1:n - 1]).sum(axis=1)
    return s * (2.0 / (n - 1))


def _dct2(x, norm):  # dct.rs:523-559
    n = x.size
    if n == 0:
        raise OracleError("ValueError", "Input array cannot be empty")
    k = np.arange(n)[:, None]
    i = np.arange(n)[None, :]
    r = (np.cos(PI * (i + 0.5) * k / n) * x[None, :]).sum(axis=1)
    if norm == "ortho":
        r *= np.sqrt(2.0 / n)
        r[0] *= 1.0 / np.sqrt(2.0)
    return r


def _idct2(x, norm):  # dct.rs:563-601
    n = x.size
    if n == 0:
        raise OracleError("ValueError", "Input array cannot be empty")
    inp = x.copy()
    if norm == "ortho":
        inp *= np.sqrt(n / 2.0)
        inp[0] *= np.sqrt(2.0)
    i = np.arange(n)[:, None]
    k = np.arange(1, n)[None, :]
    s = inp[0] * 0.5 + (np.cos(PI * k * (i + 0.5) / n) * inp[None, 1:]).sum(axis=1)
    return s * (2.0 / n)


def _dct3(x, norm):  # dct.rs:605-643
    n = x.size
    if n == 0:
        raise OracleError("ValueError", "Input array cannot be empty")
    inp = x.copy()
    if norm == "ortho":
        inp *= np.sqrt(n / 2.0)
        inp[0] *= 1.0 / np.sqrt(2.0)
    k = np.arange(n)[:, None]
    i = np.arange(1, n)[None, :]
    s = inp[0] * 0.5 + (np.cos(PI * i * (k + 0.5) / n) * inp[None, 1:]).sum(axis=1)
    return s * (2.0 / n)


def _idct3(x, norm):  # dct.rs:647-684
    n = x.size
    if n == 0:
        raise OracleError("ValueError", "Input array cannot be empty")
    inp = x.copy()
    if norm == "ortho":
        inp *= np.sqrt(2.0 / n)
        inp[0] *= np.sqrt(2.0)
    i = np.arange(n)[:, None]
    k = np.arange(n)[None, :]
    return (np.cos(PI * (i + 0.5) * k / n) * inp[None, :]).sum(axis=1)


def _dct4(x, norm):  # dct.rs:688-720
    n = x.size
    if n == 0:
        raise OracleError("ValueError", "Input array cannot be empty")
    k = np.arange(n)[:, None]
    i = np.arange(n)[None, :]
    r = (np.cos(PI * (i + 0.5) * (k + 0.5) / n) * x[None, :]).sum(axis=1)
    if norm == "ortho":
        r *= np.sqrt(2.0 / n)
    return r


def _idct4(x, norm):  # dct.rs:724-746
    n = x.size
    if n == 0:
        raise OracleError("ValueError", "Input array cannot be empty")
    inp = x * (np.sqrt(n / 2.0) if norm == "ortho" else 2.0 / n)
    return _dct4(inp, norm)


_DCT = {1: (_dct1, _idct1), 2: (_dct2, _idct2), 3: (_dct3, _idct3), 4: (_dct4, _idct4)}


def dct(x, dct_type: int = 2, norm: Optional[str] = None) -> np.ndarray:  # dct.rs:56-78
    return _DCT[dct_type][0](_f64(x), norm)


def idct(x, dct_type: int = 2, norm: Optional[str] = None) -> np.ndarray:  # dct.rs:115-138
    return _DCT[dct_type][1](_f64(x), norm)


def _along_axes(a, axes, fn):  # dct.rs:302-360: each listed axis in order, lane by lane
    r = np.array(a, dtype=np.float64)
    axes = list(range(r.ndim)) if axes is None else list(axes)
    for ax in axes:
        r = np.apply_along_axis(fn, ax, r)
    return r


def dctn(x, dct_type: int = 2, norm=None, axes=None):  # dct.rs:302-360
    return _along_axes(x, axes, lambda v: dct(v, dct_type, norm))


def idctn(x, dct_type: int = 2, norm=None, axes=None):  # dct.rs:373-420
    return _along_axes(x, axes, lambda v: idct(v, dct_type, norm))


def dct2(x, dct_type: int = 2, norm=None):  # dct.rs:168-206: rows, then columns
    return _along_axes(x, [1, 0], lambda v: dct(v, dct_type, norm))


def idct2(x, dct_type: int = 2, norm=None):  # dct.rs:242-280
    return _along_axes(x, [1, 0], lambda v: idct(v, dct_type, norm))


# ----------------------------------------------------------------------------- DST (dst.rs)

def _dst1(x, norm):  # dst.rs:409-446
    n = x.size
    if n < 2:
        raise OracleError("ValueError", "Input array must have at least 2 elements for DST-I")
    k = np.arange(1, n + 1)[:, None]
    m = np.arange(1, n + 1)[None, :]
    r = (np.sin(PI * k * m / (n + 1.0)) * x[None, :]).sum(axis=1)
    return r * (np.sqrt(2.0 / (n + 1.0)) if norm == "ortho" else 2.0 / np.sqrt(n + 1.0))


def _idst1(x, norm):  # dst.rs:450-480
    n = x.size
    if n < 2:
        raise OracleError("ValueError", "Input array must have at least 2 elements for IDST-I")
    return _dst1(x * (np.sqrt(n + 1.0) / 2.0), None)


def _dst2(x, norm):  # dst.rs:484-516
    n = x.size
    if n == 0:
        raise OracleError("ValueError", "Input array cannot be empty")
    k = np.arange(1, n + 1)[:, None]
    m = np.arange(n)[None, :]
    r = (np.sin(PI * k * (m + 0.5) / n) * x[None, :]).sum(axis=1)
    return r * (np.sqrt(2.0 / n) if norm == "ortho" else 1.0)


def _dst3(x, norm):  # dst.rs:549-592
    n = x.size
    if n == 0:
        raise OracleError("ValueError", "Input array cannot be empty")
    k = np.arange(n)
    s = x[n - 1] * np.where(k % 2 == 0, 1.0, -1.0)
    if n > 1:
        m = np.arange(1, n)[None, :]
        s = s + (np.sin(PI * m * (k[:, None] + 0.5) / n) * x[None, : n - 1]).sum(axis=1)
    return s * (np.sqrt(2.0 / n) / 2.0 if norm == "ortho" else 0.5)


def _idst2(x, norm):  # dst.rs:520-545
    n = x.size
    if n == 0:
        raise OracleError("ValueError", "Input array cannot be empty")
    return _dst3(x * (np.sqrt(n / 2.0) if norm == "ortho" else 1.0), None)


def _idst3(x, norm):  # dst.rs:596-626
    n = x.size
    if n == 0:
        raise OracleError("ValueError", "Input array cannot be empty")
    return _dst2(x * (np.sqrt(n / 2.0) * 2.0 if norm == "ortho" else 2.0), None)


def _dst4(x, norm):  # dst.rs:630-667
    n = x.size
    if n == 0:
        raise OracleError("ValueError", "Input array cannot be empty")
    k = np.arange(n)[:, None]
    m = np.arange(n)[None, :]
    r = (np.sin(PI * (m + 0.5) * (k + 0.5) / n) * x[None, :]).sum(axis=1)
    return r * (np.sqrt(2.0 / n) if norm == "ortho" else 2.0)


def _idst4(x, norm):  # dst.rs:671-702
    n = x.size
    if n == 0:
        raise OracleError("ValueError", "Input array cannot be empty")
    return _dst4(x * (np.sqrt(n / 2.0) if norm == "ortho" else 0.5), None)


_DST = {1: (_dst1, _idst1), 2: (_dst2, _idst2), 3: (_dst3, _idst3), 4: (_dst4, _idst4)}


def dst(x, dst_type: int = 2, norm: Optional[str] = None) -> np.ndarray:
    return _DST[dst_type][0](_f64(x), norm)


def idst(x, dst_type: int = 2, norm: Optional[str] = None) -> np.ndarray:
    return _DST[dst_type][1](_f64(x), norm)


def dstn(x, dst_type: int = 2, norm=None, axes=None):
    return _along_axes(x, axes, lambda v: dst(v, dst_type, norm))


def idstn(x, dst_type: int = 2, norm=None, axes=None):
    return _along_axes(x, axes, lambda v: idst(v, dst_type, norm))


def dst2(x, dst_type: int = 2, norm=None):
    return _along_axes(x, [1, 0], lambda v: dst(v, dst_type, norm))


def idst2(x, dst_type: int = 2, norm=None):
    return _along_axes(x, [1, 0], lambda v: idst(v, dst_type, norm))


# ----------------------------------------------------------------------------- Hartley (hartley.rs)

def dht(x) -> np.ndarray:  # hartley.rs:37-66
    v = _f64(x)
    n = v.size
    if n == 0:
        raise OracleError("ValueError", "empty array")
    f = base.fft(v.astype(np.complex128), None)  # pads to the next power of two
    return f[:n].real - f[:n].imag


def idht(h) -> np.ndarray:  # hartley.rs:92-112
    v = _f64(h)
    if v.size == 0:
        raise OracleError("ValueError", "empty array")
    return dht(v) / v.size


def dht2(x, axes=None) -> np.ndarray:  # hartley.rs:133-200
    a = np.asarray(x, dtype=np.float64)
    axes = (0, 1) if axes is None else tuple(axes)
    if axes[0] >= 2 or axes[1] >= 2:
        raise OracleError("ValueError", f"Axes out of bounds: {axes}")
    r = np.apply_along_axis(dht, 0 if axes[0] == 0 else 1, a)
    return np.apply_along_axis(dht, 1 if axes[1] == 1 else 0, r)


fht = dht  # hartley.rs:202-209


# ----------------------------------------------------------------------------- hfft / ihfft

def hfft(x, n: Optional[int] = None, norm=None) -> np.ndarray:  # complex_to_real.rs:58-135
    c = np.asarray(x, dtype=np.complex128).reshape(-1).copy()
    if c.size:
        c[0] = complex(c[0].real, 0.0)
    n_fft = c.size if n is None else n
    return base.fft(c, n_fft).real.copy()


def ihfft(x, n: Optional[int] = None, norm=None) -> np.ndarray:  # real_to_complex.rs:112-149
    v = _f64(x)
    n_fft = v.size if n is None else n
    c = np.zeros(n_fft, dtype=np.complex128)
    m = min(n_fft, v.size)
    c[:m] = v[:m]
    r = base.ifft(c, n_fft)
    out = np.empty(n_fft, dtype=np.complex128)
    if n_fft:
        out[0] = complex(r[0].real, 0.0)
        mid = (n_fft + 1) // 2
        out[1:mid] = r[1:mid]
        tail = [np.conj(r[i]) for i in range(n_fft - mid, 0, -1)]
        out[mid:] = tail
    return out


# ----------------------------------------------------------------------------- hilbert (lib.rs:437-516)

def hilbert(x) -> np.ndarray:
    v = np.asarray(x)
    v = (v.real if np.iscomplexobj(v) else v).astype(np.float64).reshape(-1)
    n = v.size
    spectrum = base.fft(v, None)
    h = np.ones(n, dtype=np.complex128)
    if n % 2 == 0:
        h[0] = 1.0
        h[n // 2] = 1.0
        h[1:n // 2] = -2.0j
        h[n // 2 + 1:] = 0.0
    else:
        h[0] = 1.0
        h[1:(n + 1) // 2] = -2.0j
        h[(n + 1) // 2:] = 0.0
    filtered = spectrum[:n] * h  # zip() stops at the shorter of the two
    return base.ifft(filtered, None)


# ----------------------------------------------------------------------------- windows (subset of window.rs)

def _general_cosine(n: int, sym: bool, a: Sequence[float]) -> np.ndarray:  # window.rs:542-568
    if n == 1:
        return np.ones(1)
    fac = 2.0 * PI / (n - 1.0) if sym else 2.0 * PI / n
    i = np.arange(n, dtype=np.float64)
    w = np.full(n, a[0], dtype=np.float64)
    for k in range(1, len(a)):
        w += (-1.0 if k % 2 == 1 else 1.0) * a[k] * np.cos(k * fac * i)
    return w


def get_window(window, n: int, sym: bool = True) -> np.ndarray:  # window.rs:107-142
    if n == 0:
        raise OracleError("ValueError", "Window length must be positive")
    if not isinstance(window, str):
        return np.asarray(window, dtype=np.float64)
    name = window.lower()
    if name in ("rectangular", "boxcar", "rect"):
        return np.ones(n)
    if name in ("hann", "hanning"):
        return _general_cosine(n, sym, [0.5, 0.5])
    if name == "hamming":
        return _general_cosine(n, sym, [0.54, 0.46])
    if name == "blackman":
        return _general_cosine(n, sym, [0.42, 0.5, 0.08])
    raise OracleError("NotImplementedError", f"window {window!r} is not restated in the oracle")


# ----------------------------------------------------------------------------- stft / spectrogram

def stft(x, window="hann", nperseg: int = 256, noverlap=None, nfft=None, fs=None, detrend=None,
         return_onesided=None, boundary=None):  # spectrogram.rs:76-310
    v = _f64(x)
    if v.size == 0:
        raise OracleError("ValueError", "Input signal is empty")
    if nperseg == 0:
        raise OracleError("ValueError", "Segment length must be positive")
    fs = 1.0 if fs is None else fs
    if fs <= 0.0:
        raise OracleError("ValueError", "Sampling frequency must be positive")
    nfft = nperseg if nfft is None else nfft
    if nfft < nperseg:
        raise OracleError("ValueError", "FFT length must be greater than or equal to segment length")
    noverlap = nperseg // 2 if noverlap is None else noverlap
    if noverlap >= nperseg:
        raise OracleError("ValueError", "Overlap must be less than segment length")
    detrend = True if detrend is None else detrend
    onesided = True if return_onesided is None else return_onesided
    win = get_window(window, nperseg, True)
    step = nperseg - noverlap
    padded = v
    if boundary == "reflect":
        padded = np.concatenate([v[:nperseg][::-1], v, v[v.size - nperseg:][::-1]])
    elif boundary in ("zeros", "constant"):
        lo = 0.0 if boundary == "zeros" else v[0]
        hi = 0.0 if boundary == "zeros" else v[-1]
        padded = np.concatenate([np.full(nperseg, lo), v, np.full(nperseg, hi)])
    num_frames = 1 + (padded.size - nperseg) // step
    freq_len = nfft // 2 + 1 if onesided else nfft
    freqs = np.arange(freq_len) * fs / nfft
    times = (np.arange(num_frames) * step + nperseg // 2) / fs
    out = np.zeros((freq_len, num_frames), dtype=np.complex128)
    for i in range(num_frames):
        seg = padded[i * step:i * step + nperseg].copy()
        if detrend:
            seg -= seg.sum() / seg.size
        seg = seg * win
        if nfft > nperseg:
            seg = np.concatenate([seg, np.zeros(nfft - nperseg)])
        f = base.fft(seg, None)
        rel = f[:freq_len] if onesided else f
        out[:rel.size, i] = rel  # the reference panics when rel is longer than freq_len
    return freqs, times, out


def spectrogram(x, fs=None, window=None, nperseg=None, noverlap=None, nfft=None, detrend=None, scaling=None,
                mode=None):  # spectrogram.rs:312-420
    fs = 1.0 if fs is None else fs
    window = "hann" if window is None else window
    nperseg = 256 if nperseg is None else nperseg
    freqs, times, z = stft(x, window, nperseg, noverlap, nfft, fs, detrend, True, None)
    win = get_window(window, nperseg, True)
    wss = float((win * win).sum())
    scaling = "density" if scaling is None else scaling
    if scaling == "density":
        sf = 1.0 / (fs * wss)
    elif scaling == "spectrum":
        sf = 1.0 / wss
    else:
        raise OracleError("ValueError", f"Unknown scaling mode: {scaling}. Use 'density' or 'spectrum'.")
    mode = "psd" if mode is None else mode
    if mode == "psd":
        r = (z.real ** 2 + z.imag ** 2) * sf
    elif mode == "magnitude":
        r = np.abs(z) * np.sqrt(sf)
    elif mode in ("angle", "phase"):
        r = np.angle(z)
        if mode == "angle":
            r = r * 180.0 / PI
    else:
        raise OracleError("ValueError", f"Unknown mode: {mode}. Use 'psd', 'magnitude', 'angle', or 'phase'.")
    return freqs, times, r


# ----------------------------------------------------------------------------- memory_efficient.rs / ndim_optimized.rs

def fft_inplace(inp: np.ndarray, out: np.ndarray, inverse: bool = False, normalize: bool = False) -> int:
    """memory_efficient.rs:89-190 with simd_support_available() == true (x86_64 / aarch64)."""
    n = inp.size
    if n == 0:
        raise OracleError("ValueError", "Input array is empty")
    if out.size < n:
        raise OracleError("ValueError", f"Output buffer is too small: got {out.size}, need {n}")
    if n >= 32:
        # fft_adaptive / ifft_adaptive ignore the 1-D norm (simd_fft.rs:37-60): fft(x, None) / ifft(x, None)
        r = base.ifft(inp, None) if inverse else base.fft(inp, None)
        if r.size > n:
            raise OracleError("ValueError", "index out of bounds in the reference")
    else:
        r = base._process(np.asarray(inp, dtype=np.complex128).copy(), inverse) * (1.0 / n if normalize else 1.0)
    inp[:n] = r[:n]
    out[:n] = r[:n]
    return n


def fft2_efficient(x, shape=None, inverse: bool = False, normalize: bool = False) -> np.ndarray:  # :243-397
    a = np.asarray(x)
    r, c = a.shape if shape is None else shape
    if r == 0 or c == 0:
        raise OracleError("ValueError", "Output dimensions must be positive")
    buf = np.zeros((r, c), dtype=np.complex128)
    buf[:min(r, a.shape[0]), :min(c, a.shape[1])] = a[:min(r, a.shape[0]), :min(c, a.shape[1])]
    for i in range(r):
        buf[i, :] = base._process(buf[i, :].copy(), inverse)
    for j in range(c):
        buf[:, j] = base._process(buf[:, j].copy(), inverse)
    return buf * (1.0 / (r * c) if normalize else 1.0)


def fft_streaming(x, n=None, inverse: bool = False, chunk_size=None) -> np.ndarray:  # :401-580
    a = np.asarray(x).reshape(-1).astype(np.complex128)
    L = a.size
    n_val = L if n is None else n
    chunk = chunk_size if chunk_size is not None else (1_048_576 if L > 1_000_000 else (65_536 if L > 100_000 else L))
    if L <= chunk or n_val <= chunk:
        c = np.zeros(n_val, dtype=np.complex128)
        m = min(n_val, L)
        c[:m] = a[:m]
        return base._process(c, inverse) * (1.0 / n_val if inverse else 1.0)
    chunk = max(chunk, 1)
    res = []
    for start in range(0, n_val, chunk):
        end = min(start + chunk, n_val)
        c = np.zeros(end - start, dtype=np.complex128)
        if start < L:
            e2 = min(end, L)
            c[:e2 - start] = a[start:e2]
        res.append(base._process(c, inverse) * (1.0 / (end - start) if inverse else 1.0))
    r = np.concatenate(res)
    if inverse:
        r = r * ((1.0 / n_val) / (1.0 / chunk))
    return r


def fftn_optimized(x, shape=None, axes=None) -> np.ndarray:  # ndim_optimized.rs:17-58
    r = np.asarray(x, dtype=np.float64).astype(np.complex128)
    axes = list(range(r.ndim)) if axes is None else list(axes)
    for ax in axes:
        if ax >= r.ndim:
            raise OracleError("ValueError", f"Axis {ax} is out of bounds for array with {r.ndim} dimensions")
    order = sorted(axes, key=lambda ax: int(np.prod(r.shape[ax + 1:])))
    for ax in order:
        n = r.shape[ax]
        r = np.apply_along_axis(lambda lane: base.fft(lane, None)[:n], ax, r)
    return r


# ----------------------------------------------------------------------------- czt.rs (what the file sets up; its FFTs are stubs)

def czt(x, m=None, w=None, a=None) -> np.ndarray:
    """Direct evaluation X[k] = sum_j x[j] a^-j w^(jk) along the last axis (czt.rs:45-275 describes the fast form of this sum)."""
    arr = np.asarray(x, dtype=np.complex128)
    n = arr.shape[-1]
    m = n if m is None else m
    a = 1.0 + 0j if a is None else complex(a)
    j = np.arange(n)[:, None].astype(np.float64)
    k = np.arange(m)[None, :].astype(np.float64)
    if w is None:
        kern = np.exp(-2j * np.pi * ((np.arange(n)[:, None] * np.arange(m)[None, :]) % m) / m)
    else:
        kern = np.power(complex(w), j * k)
    return (arr * np.power(a, -np.arange(n, dtype=np.float64))) @ kern
