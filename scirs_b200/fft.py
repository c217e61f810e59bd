"""Host-side mirror of the scirs2-fft free functions on the hot path.

Same names, argument meaning and error behaviour as the reference
(scirs2-fft/src/fft/algorithms.rs, scirs2-fft/src/rfft.rs, strided_fft.rs);
every call goes through the C ABI of libscirs2_fft_cuda.so with host buffers.
numpy is used only to hold the caller's arrays — no arithmetic happens here.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .error import check

_DTYPES = {
    np.dtype(np.float32): _lib.SFC_F32,
    np.dtype(np.float64): _lib.SFC_F64,
    np.dtype(np.complex64): _lib.SFC_C64,
    np.dtype(np.complex128): _lib.SFC_C128,
}


def _prep(x) -> Tuple[np.ndarray, int]:
    """C-contiguous array in one of the four boundary dtypes (other numeric types are
    widened to f64 exactly as `NumCast` does, fft/algorithms.rs:76-79)."""
    a = np.asarray(x)
    if a.dtype not in _DTYPES:
        if np.iscomplexobj(a):
            a = a.astype(np.complex128)
        else:
            a = a.astype(np.float64)
    a = np.ascontiguousarray(a)
    return a, _DTYPES[a.dtype]


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _i64arr(v: Optional[Sequence[int]]):
    if v is None:
        return None
    return (C.c_int64 * len(v))(*[int(t) for t in v])


def _norm(norm: Optional[str]):
    return None if norm is None else norm.encode()


def _next_pow2(n: int) -> int:
    p = 1
    while p < n:
        p <<= 1
    return p


# ------------------------------------------------------------------ 1-D


def fft(x, n: Optional[int] = None) -> np.ndarray:
    """`fft(&[T], Option<usize>)` — fft/algorithms.rs:131-176 (n=None pads to the next power of two)."""
    lib = _lib.load()
    a, dt = _prep(x)
    a = a.reshape(-1)
    cap = max(int(n) if n is not None else _next_pow2(max(a.size, 1)), 1)
    out = np.empty(cap, dtype=np.complex128)
    out_len = C.c_int64(0)
    check(lib.sfc_fft(_ptr(a), a.size, dt, -1 if n is None else int(n), _ptr(out), cap, C.byref(out_len)))
    return out[: out_len.value]


def ifft(x, n: Optional[int] = None) -> np.ndarray:
    """`ifft` — fft/algorithms.rs:210-263 (1/n; truncated to len when n=None padded)."""
    lib = _lib.load()
    a, dt = _prep(x)
    a = a.reshape(-1)
    cap = max(int(n) if n is not None else _next_pow2(max(a.size, 1)), 1)
    out = np.empty(cap, dtype=np.complex128)
    out_len = C.c_int64(0)
    check(lib.sfc_ifft(_ptr(a), a.size, dt, -1 if n is None else int(n), _ptr(out), cap, C.byref(out_len)))
    return out[: out_len.value]


def rfft(x, n: Optional[int] = None) -> np.ndarray:
    """`rfft` — rfft.rs:39-59 (first n/2+1 bins; n=None means len, no padding)."""
    lib = _lib.load()
    a, dt = _prep(x)
    a = a.reshape(-1)
    n_val = int(n) if n is not None else a.size
    cap = max(n_val // 2 + 1, 1)
    out = np.empty(cap, dtype=np.complex128)
    out_len = C.c_int64(0)
    check(lib.sfc_rfft(_ptr(a), a.size, dt, -1 if n is None else int(n), _ptr(out), cap, C.byref(out_len)))
    return out[: out_len.value]


def irfft(x, n: Optional[int] = None) -> np.ndarray:
    """`irfft` — rfft.rs:92-178 (n=None means 2*(len-1))."""
    lib = _lib.load()
    a, dt = _prep(x)
    a = a.reshape(-1)
    n_out = int(n) if n is not None else 2 * (a.size - 1)
    cap = max(n_out, 1)
    out = np.empty(cap, dtype=np.float64)
    out_len = C.c_int64(0)
    check(lib.sfc_irfft(_ptr(a), a.size, dt, -1 if n is None else int(n), _ptr(out), cap, C.byref(out_len)))
    return out[: out_len.value]


# ------------------------------------------------------------------ 2-D


def _two(x):
    a, dt = _prep(x)
    if a.ndim != 2:
        from .error import DimensionError

        raise DimensionError("expected a 2-D array")
    return a, dt


def _fft2(fn_name, x, shape, axes, norm):
    lib = _lib.load()
    a, dt = _two(x)
    osh = tuple(int(s) for s in shape) if shape is not None else a.shape
    out = np.empty(max(osh[0], 1) * max(osh[1], 1), dtype=np.complex128)
    oshape = (C.c_int64 * 2)()
    ax = None if axes is None else (C.c_int32 * 2)(int(axes[0]), int(axes[1]))
    check(
        getattr(lib, fn_name)(
            _ptr(a), a.shape[0], a.shape[1], dt, _i64arr(None if shape is None else osh), ax, _norm(norm), _ptr(out),
            out.size, oshape,
        )
    )
    return out[: oshape[0] * oshape[1]].reshape(oshape[0], oshape[1])


def fft2(x, shape=None, axes=None, norm: Optional[str] = None) -> np.ndarray:
    """`fft2` — fft/algorithms.rs:293-401 (axes validated then ignored; forward "backward" scales 1/N)."""
    return _fft2("sfc_fft2", x, shape, axes, norm)


def ifft2(x, shape=None, axes=None, norm: Optional[str] = None) -> np.ndarray:
    """`ifft2` — fft/algorithms.rs:439-541."""
    return _fft2("sfc_ifft2", x, shape, axes, norm)


def fft2_parallel(x, shape=None, axes=None, norm: Optional[str] = None, workers: Optional[int] = None) -> np.ndarray:
    """`fft2_parallel` — fft/planning.rs:48-201: same result as fft2; `workers` is advisory on the GPU."""
    return fft2(x, shape, axes, norm)


def ifft2_parallel(x, shape=None, axes=None, norm: Optional[str] = None, workers: Optional[int] = None) -> np.ndarray:
    """`ifft2_parallel` — fft/planning.rs:234-387."""
    return ifft2(x, shape, axes, norm)


def rfft2(x, shape=None, axes=None, norm: Optional[str] = None) -> np.ndarray:
    """`rfft2` — rfft.rs:212-232 (keeps the first n_rows/2+1 ROWS; axes and norm ignored)."""
    lib = _lib.load()
    a, dt = _two(x)
    osh = tuple(int(s) for s in shape) if shape is not None else a.shape
    out = np.empty((max(osh[0], 1) // 2 + 1) * max(osh[1], 1), dtype=np.complex128)
    oshape = (C.c_int64 * 2)()
    check(lib.sfc_rfft2(_ptr(a), a.shape[0], a.shape[1], dt, _i64arr(None if shape is None else osh), _ptr(out),
                        out.size, oshape))
    return out[: oshape[0] * oshape[1]].reshape(oshape[0], oshape[1])


def irfft2(x, shape=None, axes=None, norm: Optional[str] = None) -> np.ndarray:
    """`irfft2` — rfft.rs:274-355 (keeps the reference's (N0out*N1out)/(N0in*N1in) factor)."""
    lib = _lib.load()
    a, dt = _two(x)
    osh = tuple(int(s) for s in shape) if shape is not None else (2 * (a.shape[0] - 1), a.shape[1])
    out = np.empty(max(osh[0], 1) * max(osh[1], 1), dtype=np.float64)
    oshape = (C.c_int64 * 2)()
    check(lib.sfc_irfft2(_ptr(a), a.shape[0], a.shape[1], dt, _i64arr(None if shape is None else osh), _ptr(out),
                         out.size, oshape))
    return out[: oshape[0] * oshape[1]].reshape(oshape[0], oshape[1])


# ------------------------------------------------------------------ N-D


def _fftn(fn_name, x, shape, axes, norm):
    lib = _lib.load()
    a, dt = _prep(x)
    nd = a.ndim
    if shape is not None and len(shape) != nd:
        from .error import ValueError_

        # fft/algorithms.rs:594-598
        raise ValueError_("Output shape must have the same number of dimensions as input")
    osh = [int(s) for s in shape] if shape is not None else list(a.shape)
    total = 1
    for s in osh:
        total *= max(s, 1)
    out = np.empty(total, dtype=np.complex128)
    oshape = (C.c_int64 * max(nd, 1))()
    check(
        getattr(lib, fn_name)(
            _ptr(a), nd, _i64arr(a.shape), dt, _i64arr(None if shape is None else osh), _i64arr(axes),
            0 if axes is None else len(axes), _norm(norm), _ptr(out), out.size, oshape,
        )
    )
    res_shape = tuple(oshape[i] for i in range(nd))
    return out[: int(np.prod(res_shape))].reshape(res_shape)


def fftn(x, shape=None, axes=None, norm: Optional[str] = None, overwrite_x=None, workers=None) -> np.ndarray:
    """`fftn` — fft/algorithms.rs:576-706 (forward scale uses the product of ALL dims)."""
    return _fftn("sfc_fftn", x, shape, axes, norm)


def ifftn(x, shape=None, axes=None, norm: Optional[str] = None, overwrite_x=None, workers=None) -> np.ndarray:
    """`ifftn` — fft/algorithms.rs:757-890."""
    return _fftn("sfc_ifftn", x, shape, axes, norm)


def rfftn(x, shape=None, axes=None, norm: Optional[str] = None, overwrite_x=None, workers=None) -> np.ndarray:
    """`rfftn` — rfft.rs:472-525 (last listed axis cut to n/2+1 only when shape is None)."""
    lib = _lib.load()
    a, dt = _prep(x)
    nd = a.ndim
    if shape is not None and len(shape) != nd:
        from .error import ValueError_

        raise ValueError_("Output shape must have the same number of dimensions as input")
    osh = [int(s) for s in shape] if shape is not None else list(a.shape)
    total = 1
    for s in osh:
        total *= max(s, 1)
    out = np.empty(total, dtype=np.complex128)
    oshape = (C.c_int64 * max(nd, 1))()
    check(
        lib.sfc_rfftn(
            _ptr(a), nd, _i64arr(a.shape), dt, _i64arr(None if shape is None else osh), _i64arr(axes),
            0 if axes is None else len(axes), _norm(norm), _ptr(out), out.size, oshape,
        )
    )
    res_shape = tuple(oshape[i] for i in range(nd))
    return out[: int(np.prod(res_shape))].reshape(res_shape)


def irfftn(x, shape=None, axes=None, norm: Optional[str] = None, overwrite_x=None, workers=None) -> np.ndarray:
    """`irfftn` — rfft.rs:621-725 (Hermitian reconstruction through all axes, ifftn, real part)."""
    lib = _lib.load()
    a, dt = _prep(x)
    nd = a.ndim
    ax = list(range(nd)) if axes is None else [int(t) for t in axes]
    # capacity: resolve the output shape the same way the library will
    if shape is not None:
        if len(shape) == nd:
            osh = [int(s) for s in shape]
        elif len(shape) == len(ax):
            osh = list(a.shape)
            for i, t in enumerate(ax):
                if 0 <= t < nd:
                    osh[t] = int(shape[i])
        else:
            osh = list(a.shape)
    else:
        osh = list(a.shape)
        last = ax[-1] if ax else nd - 1
        if 0 <= last < nd:
            osh[last] = 2 * (osh[last] - 1)
    total = 1
    for s in osh:
        total *= max(s, 1)
    out = np.empty(total, dtype=np.float64)
    oshape = (C.c_int64 * max(nd, 1))()
    check(
        lib.sfc_irfftn(
            _ptr(a), nd, _i64arr(a.shape), dt, _i64arr(shape), 0 if shape is None else len(shape), _i64arr(axes),
            0 if axes is None else len(axes), _norm(norm), _ptr(out), out.size, oshape,
        )
    )
    res_shape = tuple(oshape[i] for i in range(nd))
    return out[: int(np.prod(res_shape))].reshape(res_shape)


# ------------------------------------------------------------------ strided (strided_fft.rs)


def _strided(x, axis, inverse):
    lib = _lib.load()
    a, dt = _prep(x)
    out = np.empty(a.shape, dtype=np.complex128)
    check(lib.sfc_fft_strided(_ptr(a), a.ndim, _i64arr(a.shape), dt, int(axis), 1 if inverse else 0, _ptr(out),
                              out.size))
    return out


def fft_strided(x, axis: int) -> np.ndarray:
    """`fft_strided` — strided_fft.rs:16-49 (real input along one axis)."""
    return _strided(x, axis, False)


def fft_strided_complex(x, axis: int) -> np.ndarray:
    """`fft_strided_complex` — strided_fft.rs:93-125."""
    return _strided(x, axis, False)


def ifft_strided(x, axis: int) -> np.ndarray:
    """`ifft_strided` — strided_fft.rs:166-239 (scaled by 1/len(axis))."""
    return _strided(x, axis, True)


# ------------------------------------------------------------------ aliases (simd_fft.rs, simd_rfft.rs)


def fft_simd(x, n=None, norm=None):
    """simd_fft.rs:37-50 — delegates to fft; `norm` ignored."""
    return fft(x, n)


def ifft_simd(x, n=None, norm=None):
    return ifft(x, n)


fft_adaptive = fft_simd
ifft_adaptive = ifft_simd


def fft2_simd(x, shape=None, norm=None):
    """simd_fft.rs:62-97 — delegates to fft2 on the array's own 2-D shape."""
    return fft2(x, shape, None, norm)


fft2_adaptive = fft2_simd


def fftn_simd(x, shape=None, axes=None, norm=None):
    return fftn(x, shape, axes, norm)


fftn_adaptive = fftn_simd


def ifft2_simd(*a, **k):
    """simd_fft.rs:99-111 — the reference returns NotImplementedError here."""
    from .error import NotImplementedError_

    raise NotImplementedError_("2D inverse FFT with SIMD not yet implemented")


def ifftn_simd(*a, **k):
    from .error import NotImplementedError_

    raise NotImplementedError_("N-dimensional inverse FFT with SIMD not yet implemented")


def rfft_simd(x, n=None, norm=None):
    """simd_rfft.rs:44-59 — delegates to rfft; `norm` ignored."""
    return rfft(x, n)


def irfft_simd(x, n=None, norm=None):
    return irfft(x, n)


rfft_adaptive = rfft_simd
irfft_adaptive = irfft_simd


# ------------------------------------------------------------------ batched real transforms (f32 / f64 compute)


def rfft_batch(x, prec: Optional[str] = None) -> np.ndarray:
    """[batch, n] real -> [batch, n/2+1] complex, computed in the array's own precision
    (f32 stays f32: BASELINE config 2b; the reference itself only has an f64 path)."""
    lib = _lib.load()
    a = np.ascontiguousarray(x)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)
    if a.ndim != 2:
        from .error import DimensionError

        raise DimensionError("expected a [batch, n] array")
    p = _lib.SFC_PREC_F64 if a.dtype == np.float64 else _lib.SFC_PREC_F32
    out = np.empty((a.shape[0], a.shape[1] // 2 + 1), dtype=np.complex128 if p else np.complex64)
    check(lib.sfc_rfft_batch(_ptr(a), a.shape[0], a.shape[1], p, _ptr(out)))
    return out


def irfft_batch(x, n: int) -> np.ndarray:
    """[batch, n/2+1] complex -> [batch, n] real, 1/n normalised."""
    lib = _lib.load()
    a = np.ascontiguousarray(x)
    if a.dtype not in (np.complex64, np.complex128):
        a = a.astype(np.complex128)
    if a.ndim != 2 or a.shape[1] != n // 2 + 1:
        from .error import DimensionError

        raise DimensionError("expected a [batch, n/2+1] array")
    p = _lib.SFC_PREC_F64 if a.dtype == np.complex128 else _lib.SFC_PREC_F32
    out = np.empty((a.shape[0], n), dtype=np.float64 if p else np.float32)
    check(lib.sfc_irfft_batch(_ptr(a), a.shape[0], int(n), p, _ptr(out)))
    return out
