"""Host-side mirror of the in-crate consumers of the FFT hot path.

Same names, argument meaning and error behaviour as the reference modules
(scirs2-fft/src/dct.rs, dst.rs, hartley.rs, hfft/*.rs, lib.rs `hilbert`, spectrogram.rs); every call
goes through the C ABI (include/scirs2_fft_cuda.h, "consumers of the hot path").  numpy only holds the
caller's arrays; window samples (an O(nperseg) table, window.rs) are the one thing evaluated here.
"""
from __future__ import annotations

import ctypes as C
import enum
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .error import check, ValueError_, NotImplementedError_
from .fft import _ptr, _prep


class DCTType(enum.IntEnum):
    """dct.rs:13-23"""
    Type1 = 1
    Type2 = 2
    Type3 = 3
    Type4 = 4


class DSTType(enum.IntEnum):
    """dst.rs:13-22"""
    Type1 = 1
    Type2 = 2
    Type3 = 3
    Type4 = 4


def _real(x) -> np.ndarray:
    """`NumCast` widening of any real input to f64 (dct.rs:61-67)."""
    a = np.asarray(x)
    if np.iscomplexobj(a):
        raise ValueError_(f"Could not convert {a.dtype} to f64")
    return np.ascontiguousarray(a, dtype=np.float64)


def _trig(fn_name: str, x, ttype, inverse: bool, norm: Optional[str], axes: Optional[Sequence[int]]) -> np.ndarray:
    lib = _lib.load()
    a = _real(x)
    shape = (C.c_int64 * max(a.ndim, 1))(*([int(s) for s in a.shape] or [0]))
    ax = None if axes is None else (C.c_int32 * max(len(axes), 1))(*[int(v) for v in axes])
    out = np.empty(a.shape, dtype=np.float64)
    check(getattr(lib, fn_name)(_ptr(a), a.ndim, shape, ax, 0 if axes is None else len(axes), int(ttype), int(inverse),
                                None if norm is None else norm.encode(), _ptr(out)))
    return out


# ------------------------------------------------------------------ DCT (dct.rs:56-420)

def dct(x, dct_type: Optional[DCTType] = None, norm: Optional[str] = None) -> np.ndarray:
    """`dct(&[T], Option<DCTType>, Option<&str>)` — dct.rs:56-78 (default Type2)."""
    return _trig("sfc_dct", np.asarray(x).reshape(-1), dct_type or DCTType.Type2, False, norm, [0])


def idct(x, dct_type: Optional[DCTType] = None, norm: Optional[str] = None) -> np.ndarray:
    """dct.rs:115-138"""
    return _trig("sfc_dct", np.asarray(x).reshape(-1), dct_type or DCTType.Type2, True, norm, [0])


def dct2(x, dct_type: Optional[DCTType] = None, norm: Optional[str] = None) -> np.ndarray:
    """dct.rs:168-206: rows first, then columns."""
    return _trig("sfc_dct", x, dct_type or DCTType.Type2, False, norm, [1, 0])


def idct2(x, dct_type: Optional[DCTType] = None, norm: Optional[str] = None) -> np.ndarray:
    """dct.rs:242-280"""
    return _trig("sfc_dct", x, dct_type or DCTType.Type2, True, norm, [1, 0])


def dctn(x, dct_type: Optional[DCTType] = None, norm: Optional[str] = None, axes: Optional[Sequence[int]] = None):
    """dct.rs:302-360 (axes=None: every axis, in order)."""
    return _trig("sfc_dct", x, dct_type or DCTType.Type2, False, norm, axes)


def idctn(x, dct_type: Optional[DCTType] = None, norm: Optional[str] = None, axes: Optional[Sequence[int]] = None):
    """dct.rs:373-420"""
    return _trig("sfc_dct", x, dct_type or DCTType.Type2, True, norm, axes)


# ------------------------------------------------------------------ DST (dst.rs:48-405)

def dst(x, dst_type: Optional[DSTType] = None, norm: Optional[str] = None) -> np.ndarray:
    return _trig("sfc_dst", np.asarray(x).reshape(-1), dst_type or DSTType.Type2, False, norm, [0])


def idst(x, dst_type: Optional[DSTType] = None, norm: Optional[str] = None) -> np.ndarray:
    return _trig("sfc_dst", np.asarray(x).reshape(-1), dst_type or DSTType.Type2, True, norm, [0])


def dst2(x, dst_type: Optional[DSTType] = None, norm: Optional[str] = None) -> np.ndarray:
    return _trig("sfc_dst", x, dst_type or DSTType.Type2, False, norm, [1, 0])


def idst2(x, dst_type: Optional[DSTType] = None, norm: Optional[str] = None) -> np.ndarray:
    return _trig("sfc_dst", x, dst_type or DSTType.Type2, True, norm, [1, 0])


def dstn(x, dst_type: Optional[DSTType] = None, norm: Optional[str] = None, axes: Optional[Sequence[int]] = None):
    return _trig("sfc_dst", x, dst_type or DSTType.Type2, False, norm, axes)


def idstn(x, dst_type: Optional[DSTType] = None, norm: Optional[str] = None, axes: Optional[Sequence[int]] = None):
    return _trig("sfc_dst", x, dst_type or DSTType.Type2, True, norm, axes)


# ------------------------------------------------------------------ Hartley (hartley.rs)

def dht(x) -> np.ndarray:
    """hartley.rs:37-66 (input of any shape is flattened)."""
    lib = _lib.load()
    a = _real(x).reshape(-1)
    out = np.empty(a.size, dtype=np.float64)
    check(lib.sfc_dht(_ptr(a), a.size, _ptr(out)))
    return out


def idht(h) -> np.ndarray:
    """hartley.rs:92-112"""
    lib = _lib.load()
    a = _real(h).reshape(-1)
    out = np.empty(a.size, dtype=np.float64)
    check(lib.sfc_idht(_ptr(a), a.size, _ptr(out)))
    return out


def dht2(x, axes: Optional[Tuple[int, int]] = None) -> np.ndarray:
    """hartley.rs:133-200"""
    lib = _lib.load()
    a = _real(x)
    if a.ndim != 2:
        raise ValueError_("dht2 needs a 2-D array")
    ax = (0, 1) if axes is None else tuple(int(v) for v in axes)
    out = np.empty(a.shape, dtype=np.float64)
    check(lib.sfc_dht2(_ptr(a), a.shape[0], a.shape[1], ax[0], ax[1], _ptr(out)))
    return out


def fht(x) -> np.ndarray:
    """hartley.rs:202-209: alias of dht."""
    return dht(x)


# ------------------------------------------------------------------ hfft / ihfft

def hfft(x, n: Optional[int] = None, norm: Optional[str] = None) -> np.ndarray:
    """hfft/complex_to_real.rs:58-135 (`norm` is ignored there too)."""
    lib = _lib.load()
    a, dt = _prep(x)
    a = a.reshape(-1)
    cap = max(int(n) if n is not None else a.size, 1)
    out = np.empty(cap, dtype=np.float64)
    out_len = C.c_int64(0)
    check(lib.sfc_hfft(_ptr(a), a.size, dt, -1 if n is None else int(n), _ptr(out), cap, C.byref(out_len)))
    return out[: out_len.value]


def ihfft(x, n: Optional[int] = None, norm: Optional[str] = None) -> np.ndarray:
    """hfft/real_to_complex.rs:49-149"""
    lib = _lib.load()
    a = _real(x).reshape(-1)
    cap = max(int(n) if n is not None else a.size, 1)
    out = np.empty(cap, dtype=np.complex128)
    out_len = C.c_int64(0)
    check(lib.sfc_ihfft(_ptr(a), a.size, -1 if n is None else int(n), _ptr(out), cap, C.byref(out_len)))
    return out[: out_len.value]


# ------------------------------------------------------------------ hilbert (lib.rs:437-516)

def hilbert(x) -> np.ndarray:
    """Analytic signal; complex input contributes its real part only (lib.rs:455-463)."""
    lib = _lib.load()
    a = np.asarray(x)
    a = np.ascontiguousarray((a.real if np.iscomplexobj(a) else a), dtype=np.float64).reshape(-1)
    out = np.empty(a.size, dtype=np.complex128)
    check(lib.sfc_hilbert(_ptr(a), a.size, _ptr(out)))
    return out


# ------------------------------------------------------------------ windows (window.rs, the STFT's table)

def _general_cosine(n: int, sym: bool, a: Sequence[float]) -> np.ndarray:
    """window.rs:542-568"""
    if n == 1:
        return np.ones(1)
    fac = 2.0 * np.pi / (n - 1.0) if sym else 2.0 * np.pi / n
    i = np.arange(n, dtype=np.float64)
    w = np.full(n, a[0], dtype=np.float64)
    for k in range(1, len(a)):
        w += (-1.0 if k % 2 == 1 else 1.0) * a[k] * np.cos(k * fac * i)
    return w


def get_window(window, n: int, sym: bool = True) -> np.ndarray:
    """window.rs:107-142 for the cosine-sum family and rectangular; any other window can be passed to
    stft / spectrogram as an array of nperseg samples."""
    if n == 0:
        raise ValueError_("Window length must be positive")
    if not isinstance(window, str):
        w = np.ascontiguousarray(window, dtype=np.float64).reshape(-1)
        if w.size != n:
            raise ValueError_("window array must hold nperseg samples")
        return w
    name = window.lower()
    if name in ("rectangular", "boxcar", "rect"):
        return np.ones(n)
    if name in ("hann", "hanning"):
        return _general_cosine(n, sym, [0.5, 0.5])
    if name == "hamming":
        return _general_cosine(n, sym, [0.54, 0.46])
    if name == "blackman":
        return _general_cosine(n, sym, [0.42, 0.5, 0.08])
    raise NotImplementedError_(f"window {window!r}: pass its samples as an array")


# ------------------------------------------------------------------ stft / spectrogram (spectrogram.rs)

_BOUNDARY = {None: 0, "reflect": 1, "zeros": 2, "constant": 3}
_MODES = {"complex": 0, "psd": 1, "magnitude": 2, "phase": 3, "angle": 4}


def _stft_call(x, win, nperseg, noverlap, nfft, detrend, onesided, boundary, mode, scale):
    lib = _lib.load()
    a = _real(x).reshape(-1)
    nfft_v = nperseg if nfft is None else int(nfft)
    nov = nperseg // 2 if noverlap is None else int(noverlap)
    if a.size == 0:
        raise ValueError_("Input signal is empty")
    if nperseg == 0:
        raise ValueError_("Segment length must be positive")
    step = max(nperseg - nov, 1)
    padded = a.size + (2 * nperseg if boundary is not None and boundary in _BOUNDARY else 0)
    frames = max(1 + (padded - nperseg) // step, 0) if padded >= nperseg else 0
    freq_len = nfft_v // 2 + 1 if onesided else nfft_v
    out = np.empty((max(freq_len, 1), max(frames, 1)), dtype=np.complex128 if mode == 0 else np.float64)
    fl, fr = C.c_int64(0), C.c_int64(0)
    check(lib.sfc_stft(_ptr(a), a.size, _ptr(win), int(nperseg), nov, nfft_v, int(bool(detrend)), int(bool(onesided)),
                       _BOUNDARY.get(boundary, 0), mode, float(scale), _ptr(out), out.size, C.byref(fl), C.byref(fr)))
    return out.reshape(-1)[: fl.value * fr.value].reshape(fl.value, fr.value), nfft_v, step


def stft(x, window="hann", nperseg: int = 256, noverlap: Optional[int] = None, nfft: Optional[int] = None,
         fs: Optional[float] = None, detrend: Optional[bool] = None, return_onesided: Optional[bool] = None,
         boundary: Optional[str] = None):
    """`stft` — spectrogram.rs:76-310: (frequencies, times, Zxx[freq][frame])."""
    fs = 1.0 if fs is None else fs
    if fs <= 0.0:
        raise ValueError_("Sampling frequency must be positive")
    win = get_window(window, nperseg, True) if nperseg > 0 else np.ones(1)
    onesided = True if return_onesided is None else return_onesided
    z, nfft_v, step = _stft_call(x, win, nperseg, noverlap, nfft, True if detrend is None else detrend, onesided,
                                 boundary, 0, 1.0)
    freqs = np.arange(z.shape[0]) * fs / nfft_v
    times = (np.arange(z.shape[1]) * step + nperseg // 2) / fs
    return freqs, times, z


def spectrogram(x, fs: Optional[float] = None, window=None, nperseg: Optional[int] = None,
                noverlap: Optional[int] = None, nfft: Optional[int] = None, detrend: Optional[bool] = None,
                scaling: Optional[str] = None, mode: Optional[str] = None):
    """`spectrogram` — spectrogram.rs:312-420."""
    fs = 1.0 if fs is None else fs
    if fs <= 0.0:
        raise ValueError_("Sampling frequency must be positive")
    window = "hann" if window is None else window
    nperseg = 256 if nperseg is None else nperseg
    win = get_window(window, nperseg, True)
    wss = float((win * win).sum())
    scaling = "density" if scaling is None else scaling
    if scaling == "density":
        sf = 1.0 / (fs * wss)
    elif scaling == "spectrum":
        sf = 1.0 / wss
    else:
        raise ValueError_(f"Unknown scaling mode: {scaling}. Use 'density' or 'spectrum'.")
    mode = "psd" if mode is None else mode
    if mode not in ("psd", "magnitude", "angle", "phase"):
        raise ValueError_(f"Unknown mode: {mode}. Use 'psd', 'magnitude', 'angle', or 'phase'.")
    r, nfft_v, step = _stft_call(x, win, nperseg, noverlap, nfft, True if detrend is None else detrend, True, None,
                                 _MODES[mode], sf)
    freqs = np.arange(r.shape[0]) * fs / nfft_v
    times = (np.arange(r.shape[1]) * step + nperseg // 2) / fs
    return freqs, times, r


# ------------------------------------------------------------------ memory_efficient.rs / ndim_optimized.rs

class FftMode(enum.IntEnum):
    """memory_efficient.rs:41-46"""
    Forward = 0
    Inverse = 1


def fft_inplace(input: np.ndarray, output: np.ndarray, mode: FftMode = FftMode.Forward, normalize: bool = False) -> int:
    """memory_efficient.rs:89-190: transforms `input` (complex128, modified in place) and mirrors the result into
    `output`; returns the number of elements."""
    lib = _lib.load()
    if not (isinstance(input, np.ndarray) and input.dtype == np.complex128 and input.flags.c_contiguous):
        raise ValueError_("fft_inplace needs a contiguous complex128 input buffer")
    if not (isinstance(output, np.ndarray) and output.dtype == np.complex128 and output.flags.c_contiguous):
        raise ValueError_("fft_inplace needs a contiguous complex128 output buffer")
    rc = lib.sfc_fft_inplace(_ptr(input), input.size, _ptr(output), output.size, int(mode), int(bool(normalize)))
    check(rc)
    return input.size


def process_in_chunks(input, chunk_size: int, op):
    """memory_efficient.rs:194-226: apply `op` to consecutive chunks and concatenate."""
    a = np.asarray(input).reshape(-1)
    if a.size <= chunk_size:
        return op(a)
    chunk_size = max(int(chunk_size), 1)
    return np.concatenate([np.asarray(op(a[s:s + chunk_size])) for s in range(0, a.size, chunk_size)])


def fft2_efficient(input, shape: Optional[Tuple[int, int]] = None, mode: FftMode = FftMode.Forward,
                   normalize: bool = False) -> np.ndarray:
    """memory_efficient.rs:243-397"""
    lib = _lib.load()
    a, dt = _prep(input)
    if a.ndim != 2:
        raise ValueError_("fft2_efficient needs a 2-D array")
    r, c = (a.shape if shape is None else (int(shape[0]), int(shape[1])))
    out = np.empty((max(r, 1), max(c, 1)), dtype=np.complex128)
    check(lib.sfc_fft2_efficient(_ptr(a), a.shape[0], a.shape[1], dt, r, c, int(mode), int(bool(normalize)), _ptr(out)))
    return out


def fft_streaming(input, n: Optional[int] = None, mode: FftMode = FftMode.Forward,
                  chunk_size: Optional[int] = None) -> np.ndarray:
    """memory_efficient.rs:401-580"""
    lib = _lib.load()
    a, dt = _prep(input)
    a = a.reshape(-1)
    n_val = a.size if n is None else int(n)
    out = np.empty(max(n_val, 1), dtype=np.complex128)
    check(lib.sfc_fft_streaming(_ptr(a), a.size, dt, -1 if n is None else n_val, int(mode),
                                -1 if chunk_size is None else int(chunk_size), _ptr(out)))
    return out[:n_val]


def fftn_optimized(x, shape=None, axes: Optional[Sequence[int]] = None) -> np.ndarray:
    """ndim_optimized.rs:17-58 (`shape` is ignored there too)."""
    lib = _lib.load()
    a = _real(x)
    sh = (C.c_int64 * max(a.ndim, 1))(*([int(s) for s in a.shape] or [0]))
    ax = None if axes is None else (C.c_int32 * max(len(axes), 1))(*[int(v) for v in axes])
    out = np.empty(a.shape, dtype=np.complex128)
    check(lib.sfc_fftn_optimized(_ptr(a), a.ndim, sh, ax, 0 if axes is None else len(axes), _ptr(out)))
    return out


def fftn_memory_efficient(x, axes: Optional[Sequence[int]] = None, max_memory_gb: float = 0.0) -> np.ndarray:
    """ndim_optimized.rs:159-211: the same per-axis `fft(&lane, None)` loop as `fftn_optimized`, one axis at a time
    (`max_memory_gb` is ignored there too).  Axes longer than 2^20 take the reference's "simplified chunking"
    (:214-249: independent 65,536-point transforms of consecutive chunks, the last chunk padded to a power of two and cut
    back) — reproduced as written, because that is what a caller of the reference gets."""
    a = _real(x)
    ax = list(range(a.ndim)) if axes is None else [int(v) for v in axes]
    for v in ax:
        if v < 0 or v >= a.ndim:
            raise ValueError_(f"Axis {v} is out of bounds for array with {a.ndim} dimensions")
    if all(a.shape[v] <= 1048576 for v in ax):
        return fftn_optimized(a, None, ax)
    from .fft import fftn as _fftn

    res = a.astype(np.complex128)
    for v in ax:
        n = res.shape[v]
        if n <= 1048576:
            p = 1 << max(n - 1, 0).bit_length()
            res = np.ascontiguousarray(np.take(_fftn(res, [p if i == v else s for i, s in enumerate(res.shape)], [v]), range(n), axis=v))
            continue
        moved = np.moveaxis(res, v, -1)
        out = np.empty_like(moved)
        for s0 in range(0, n, 65536):
            e0 = min(s0 + 65536, n)
            cl = e0 - s0
            p = 1 << max(cl - 1, 0).bit_length()
            chunk = np.ascontiguousarray(moved[..., s0:e0])
            out[..., s0:e0] = _fftn(chunk, list(chunk.shape[:-1]) + [p], [chunk.ndim - 1])[..., :cl]
        res = np.ascontiguousarray(np.moveaxis(out, -1, v))
    return res


def rfftn_optimized(x, shape=None, axes: Optional[Sequence[int]] = None) -> np.ndarray:
    """ndim_optimized.rs:252-301.  The reference is a placeholder there (its real transform is computed and dropped, the
    result array stays zero); this is the transform that code sets out to compute: the complex spectrum of the real input
    over `axes` (every axis by default), in an array of the input's shape, through the same per-axis
    `fft(&lane, None)` semantics as `fftn_optimized`."""
    return fftn_optimized(x, shape, axes)
