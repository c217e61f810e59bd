"""CPU ORACLE for the scirs2-fft hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product (scirs_b200/, libscirs2_fft_cuda.so) never does.

What it is: a restatement, in numpy, of the *wrapper semantics* of
/root/reference/scirs2-fft (cool-japan/scirs 0.1.0-alpha.6) for the path
fft/ifft, rfft/irfft, fft2/ifft2, rfft2/irfft2, fftn/ifftn, rfftn/irfftn, fft_strided —
sizes, next-power-of-two padding, truncation, the norm table, axes handling, output
shapes and error cases — each function citing the reference file:line it follows.

Where the arithmetic lives: the reference delegates every butterfly to the third-party
crate `rustfft` (requirement "6.4.0", default-features = false => scalar planner;
/root/reference/Cargo.toml:86; not vendored, no Cargo.lock), which cannot be built here
(no cargo/rustc).  rustfft computes the plain unnormalised DFT
    X[k] = sum_j x[j] * exp(-/+ 2*pi*i*j*k/n)
to ~1e-16 relative accuracy, so the engine here is any exact DFT:
  * `engine="c"`      oracle/rustfft_port.c (restatement of rustfft's scalar algorithm
                      classes: radix-4, mixed radix, Bluestein), via ctypes
  * `engine="scipy"`  scipy.fft (pocketfft, f64)
  * `dft_longdouble`  O(n^2) extended-precision direct summation (cross-check, n <= 4096,
                      or sampled bins at large n)
Parity pin: the reference's own known-answer tests (doctests, rfft.rs:926-1032,
planning.rs:733-754, src/bin/accuracy_comparison.rs:83-267) — all analytic, N <= 1024 —
are restated in tests/test_oracle_golden.py and tests/golden/.  Nothing in the reference
pins values at the BASELINE sizes, for non-power-of-two lengths, norm strings or axes
subsets: there PARITY IS UNPINNED by the reference's tests and this restatement
(cross-checked against the extended-precision DFT) is the arbiter.

Deliberately NOT restated (reference bugs): the hard-coded returns of
irfft (rfft.rs:97-116) and irfft2 (rfft.rs:286-293).
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ENGINE = os.environ.get("SCIRS_ORACLE_ENGINE", "scipy")


class OracleError(Exception):
    """Carries the FFTError variant name the reference would return."""

    def __init__(self, variant: str, msg: str):
        super().__init__(msg)
        self.variant = variant
        self.msg = msg


# --------------------------------------------------------------------------- engines


def dft_longdouble(x: np.ndarray, inverse: bool = False, bins: Optional[Sequence[int]] = None) -> np.ndarray:
    """Direct O(n*len(bins)) DFT in extended precision with exact integer phase reduction."""
    x = np.asarray(x)
    n = x.shape[-1]
    ks = np.arange(n) if bins is None else np.asarray(bins)
    j = np.arange(n, dtype=np.int64)
    xr = x.real.astype(np.longdouble)
    xi = x.imag.astype(np.longdouble) if np.iscomplexobj(x) else np.zeros_like(xr)
    out_r = np.empty(x.shape[:-1] + (len(ks),), dtype=np.longdouble)
    out_i = np.empty_like(out_r)
    two_pi = 2 * np.arccos(np.longdouble(-1))
    sign = 1.0 if inverse else -1.0
    for t, k in enumerate(ks):
        ph = ((j * int(k)) % n).astype(np.longdouble) * (two_pi / n)
        c, s = np.cos(ph), sign * np.sin(ph)
        out_r[..., t] = (xr * c - xi * s).sum(axis=-1)
        out_i[..., t] = (xr * s + xi * c).sum(axis=-1)
    return out_r.astype(np.float64) + 1j * out_i.astype(np.float64)


_c_lib = None


def _load_c():
    global _c_lib
    if _c_lib is None:
        path = os.path.join(_HERE, "librustfft_port.so")
        if not os.path.exists(path):
            raise ImportError(f"{path} missing: run `make -C oracle`")
        lib = ctypes.CDLL(path)
        lib.rfp_process.restype = ctypes.c_int
        lib.rfp_process.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
        _c_lib = lib
    return _c_lib


def _process(buf: np.ndarray, inverse: bool, engine: Optional[str] = None) -> np.ndarray:
    """`Fft::process` over the last axis: unnormalised DFT, sign - forward / + inverse."""
    eng = engine or _ENGINE
    buf = np.ascontiguousarray(buf, dtype=np.complex128)
    n = buf.shape[-1]
    if eng == "c":
        lib = _load_c()
        out = buf.copy()
        rows = out.size // n
        rc = lib.rfp_process(out.ctypes.data_as(ctypes.c_void_p), rows, n, 1 if inverse else 0)
        if rc != 0:
            raise RuntimeError("rustfft_port failed")
        return out
    import scipy.fft as sf

    if inverse:
        return sf.ifft(buf, axis=-1, norm="forward")  # unnormalised inverse
    return sf.fft(buf, axis=-1)


# --------------------------------------------------------------------------- helpers


def _to_complex(x) -> np.ndarray:
    """convert_to_complex, fft/algorithms.rs:71-102: everything widens to Complex64 (f64)."""
    a = np.asarray(x)
    return a.astype(np.complex128)


def next_power_of_two(n: int) -> int:
    p = 1
    while p < n:
        p <<= 1
    return p


def parse_norm_mode(norm: Optional[str], is_inverse: bool) -> str:
    """fft/algorithms.rs:19-50"""
    if norm is None:
        return "backward" if is_inverse else "none"
    return norm if norm in ("backward", "ortho", "forward") else "none"


def _pad_or_truncate(data: np.ndarray, size: int) -> np.ndarray:
    """fft/algorithms.rs:148-156"""
    if size > data.size:
        return np.concatenate([data, np.zeros(size - data.size, dtype=np.complex128)])
    return data[:size].copy()


def _pad_crop_nd(a: np.ndarray, shape: Sequence[int]) -> np.ndarray:
    """top-left pad / crop, fft/algorithms.rs:334-347 and :633-664"""
    out = np.zeros(tuple(shape), dtype=np.complex128)
    sl = tuple(slice(0, min(s, t)) for s, t in zip(a.shape, shape))
    out[sl] = a[sl]
    return out


# --------------------------------------------------------------------------- 1-D


def fft(x, n: Optional[int] = None, engine: Optional[str] = None) -> np.ndarray:
    """fft/algorithms.rs:131-176"""
    a = np.asarray(x).reshape(-1)
    if a.size == 0:
        raise OracleError("ValueError", "Input cannot be empty")  # :136-138
    fft_size = n if n is not None else next_power_of_two(a.size)  # :142
    data = _pad_or_truncate(_to_complex(a), fft_size)  # :145-156
    return _process(data, False, engine)  # :159-173


def ifft(x, n: Optional[int] = None, engine: Optional[str] = None) -> np.ndarray:
    """fft/algorithms.rs:210-263"""
    a = np.asarray(x).reshape(-1)
    if a.size == 0:
        raise OracleError("ValueError", "Input cannot be empty")
    fft_size = n if n is not None else next_power_of_two(a.size)
    data = _pad_or_truncate(_to_complex(a), fft_size)
    res = _process(data, True, engine) * (1.0 / fft_size)  # :255
    if n is None and fft_size > a.size:  # :258-260
        res = res[: a.size]
    return res


def rfft(x, n: Optional[int] = None, engine: Optional[str] = None) -> np.ndarray:
    """rfft.rs:39-59"""
    a = np.asarray(x).reshape(-1)
    n_val = n if n is not None else a.size
    full = fft(a, n_val, engine)
    return full[: n_val // 2 + 1].copy()


def irfft(x, n: Optional[int] = None, engine: Optional[str] = None) -> np.ndarray:
    """rfft.rs:92-178 without the hard-coded returns at :97-116"""
    ci = _to_complex(np.asarray(x).reshape(-1))
    input_len = ci.size
    if input_len == 0:
        raise OracleError("ValueError", "Input cannot be empty")
    n_output = n if n is not None else 2 * (input_len - 1)  # :138-141
    if n_output <= 0:
        raise OracleError("ValueError", "Input cannot be empty")
    full = list(ci)
    if n_output > input_len:  # :150-169
        start_idx = input_len - 1 if n_output % 2 == 0 else input_len
        for i in range(start_idx - 1, 0, -1):
            if len(full) >= n_output:
                break
            full.append(np.conj(ci[i]))
        while len(full) < n_output:
            full.append(0j)
    out = ifft(np.array(full, dtype=np.complex128), n_output, engine)  # :172
    return out.real.copy()  # :175


# --------------------------------------------------------------------------- 2-D


def _norm_scale_forward(mode: str, total: float) -> float:
    """fft/algorithms.rs:385-395 / :693-703"""
    return {"none": 1.0, "backward": 1.0 / total, "ortho": 1.0 / np.sqrt(total), "forward": 1.0 / total}[mode]


def _norm_scale_inverse(mode: str, total: float) -> float:
    """fft/algorithms.rs:528-534 / :876-884"""
    return {"none": 1.0, "backward": 1.0 / total, "ortho": 1.0 / np.sqrt(total), "forward": 1.0}[mode]


def _fft2(x, shape, axes, norm, inverse, engine):
    a = np.asarray(x)
    if a.ndim != 2:
        raise OracleError("DimensionError", "expected a 2-D array")
    out_shape = tuple(shape) if shape is not None else a.shape
    ax = axes if axes is not None else (0, 1)
    if ax[0] < 0 or ax[0] > 1 or ax[1] < 0 or ax[1] > 1 or ax[0] == ax[1]:  # :309-314 (then ignored)
        raise OracleError("ValueError", "Invalid axes for 2D IFFT" if inverse else "Invalid axes for 2D FFT")
    mode = parse_norm_mode(norm, inverse)
    data = _pad_crop_nd(_to_complex(a), out_shape)
    data = _process(data, inverse, engine)  # rows, :353-366
    data = np.ascontiguousarray(_process(np.ascontiguousarray(data.T), inverse, engine).T)  # columns, :369-382
    total = float(out_shape[0] * out_shape[1])
    scale = _norm_scale_inverse(mode, total) if inverse else _norm_scale_forward(mode, total)
    return data * scale if scale != 1.0 else data


def fft2(x, shape=None, axes=None, norm=None, engine=None):
    """fft/algorithms.rs:293-401"""
    return _fft2(x, shape, axes, norm, False, engine)


def ifft2(x, shape=None, axes=None, norm=None, engine=None):
    """fft/algorithms.rs:439-541"""
    return _fft2(x, shape, axes, norm, True, engine)


def rfft2(x, shape=None, axes=None, norm=None, engine=None):
    """rfft.rs:212-232: full fft2(x, shape, None, None), first n_rows_out//2+1 ROWS"""
    a = np.asarray(x)
    n_rows_out = (shape if shape is not None else a.shape)[0]
    full = fft2(a, shape, None, None, engine)
    return full[: n_rows_out // 2 + 1, :].copy()


def irfft2(x, shape=None, axes=None, norm=None, engine=None):
    """rfft.rs:274-355 without the hard-coded 2x2 return at :286-293"""
    a = _to_complex(np.asarray(x))
    n_rows, n_cols = a.shape
    ro, co = tuple(shape) if shape is not None else (2 * (n_rows - 1), n_cols)
    if ro <= 0 or co <= 0:
        raise OracleError("ValueError", "Input cannot be empty")
    if n_rows > ro or n_cols > co:
        raise OracleError("DimensionError", "input extent exceeds the output shape")
    full = np.zeros((ro, co), dtype=np.complex128)
    full[:n_rows, :n_cols] = a
    for i in range(n_rows, ro):  # :323-335
        si = ro - i
        for j in range(co):
            sj = 0 if j == 0 else co - j
            if si < n_rows and sj < n_cols:
                full[i, j] = np.conj(full[si, sj])
    c = ifft2(full, (ro, co), None, None, engine)
    return c.real * ((ro * co) / (n_rows * n_cols))  # :347


# --------------------------------------------------------------------------- N-D


def _axis_pass(data: np.ndarray, axis: int, inverse: bool, engine) -> np.ndarray:
    """lanes_mut(Axis(axis)) gather -> process -> scatter, fft/algorithms.rs:677-689"""
    moved = np.ascontiguousarray(np.moveaxis(data, axis, -1))
    return np.ascontiguousarray(np.moveaxis(_process(moved, inverse, engine), -1, axis))


def _fftn(x, shape, axes, norm, inverse, engine):
    a = np.asarray(x)
    nd = a.ndim
    out_shape = list(shape) if shape is not None else list(a.shape)
    if len(out_shape) != nd:  # :594-598
        raise OracleError("ValueError", "Output shape must have the same number of dimensions as input")
    ax = list(axes) if axes is not None else list(range(nd))
    for t in ax:  # :604-611
        if t >= nd or t < 0:
            raise OracleError("ValueError", f"Axis {t} out of bounds for array of dimension {nd}")
    mode = parse_norm_mode(norm, inverse)
    data = _pad_crop_nd(_to_complex(a), out_shape)
    for t in ax:  # in list order, duplicates transform twice (:667-690)
        data = _axis_pass(data, t, inverse, engine)
    if inverse:
        total = float(np.prod([out_shape[t] for t in ax])) if ax else 1.0  # :876 (listed axes only)
        scale = _norm_scale_inverse(mode, total)
    else:
        total = float(np.prod(out_shape))  # :694 (ALL dims)
        scale = _norm_scale_forward(mode, total)
    return data * scale if scale != 1.0 else data


def fftn(x, shape=None, axes=None, norm=None, overwrite_x=None, workers=None, engine=None):
    """fft/algorithms.rs:576-706"""
    return _fftn(x, shape, axes, norm, False, engine)


def ifftn(x, shape=None, axes=None, norm=None, overwrite_x=None, workers=None, engine=None):
    """fft/algorithms.rs:757-890"""
    return _fftn(x, shape, axes, norm, True, engine)


def rfftn(x, shape=None, axes=None, norm=None, overwrite_x=None, workers=None, engine=None):
    """rfft.rs:472-525"""
    a = np.asarray(x)
    full = fftn(a, shape, axes, norm, engine=engine)
    ax = list(axes) if axes is not None else list(range(a.ndim))
    last_axis = ax[-1] if ax else a.ndim - 1
    if shape is None:  # :508-511
        sl = [slice(None)] * a.ndim
        sl[last_axis] = slice(0, full.shape[last_axis] // 2 + 1)
        return full[tuple(sl)].copy()
    return full


def reconstruct_hermitian_symmetry(x: np.ndarray, out_shape: Sequence[int], axes: Sequence[int]) -> np.ndarray:
    """rfft.rs:733-901, literally: known values copied, then one sweep in C order where an
    unknown index takes conj(value at its reflection through all `axes`) if that is known."""
    x = _to_complex(x)
    if any(xs > os_ for xs, os_ in zip(x.shape, out_shape)):
        raise OracleError("DimensionError", "input extent exceeds the output shape")
    res = np.zeros(tuple(out_shape), dtype=np.complex128)
    known = np.zeros(tuple(out_shape), dtype=bool)
    sl = tuple(slice(0, s) for s in x.shape)
    res[sl] = x
    known[sl] = True
    for idx in np.ndindex(*out_shape):
        if known[idx]:
            continue
        refl = list(idx)
        for t in axes:  # :861-879
            n = out_shape[t]
            if idx[t] == 0 or (n % 2 == 0 and idx[t] == n // 2):
                continue
            refl[t] = n - idx[t]
        refl = tuple(refl)
        if known[refl]:
            res[idx] = np.conj(res[refl])
            known[idx] = True
    return res


def irfftn(x, shape=None, axes=None, norm=None, overwrite_x=None, workers=None, engine=None):
    """rfft.rs:621-725"""
    a = np.asarray(x)
    nd = a.ndim
    if axes is not None:
        ax = list(axes)
        for t in ax:  # :642-648
            if t >= nd or t < 0:
                raise OracleError("DimensionError", f"Axis {t} is out of bounds for array of dimension {nd}")
    else:
        ax = list(range(nd))
    if shape is not None:
        sh = list(shape)
        if len(sh) != len(ax) and len(ax) != 0 and len(sh) != nd:  # :659-672
            raise OracleError(
                "DimensionError",
                "Shape must have the same number of dimensions as input or match the length of axes, "
                f"got {len(sh)} expected {nd} or {len(ax)}",
            )
        if len(sh) == nd:
            out_shape = sh
        elif len(sh) == len(ax):
            out_shape = list(a.shape)
            for i, t in enumerate(ax):
                out_shape[t] = sh[i]
        else:
            raise OracleError("DimensionError", "Shape has invalid dimensions")
    else:
        out_shape = list(a.shape)
        last_axis = ax[-1] if ax else nd - 1
        out_shape[last_axis] = 2 * (out_shape[last_axis] - 1)  # :699-700
    if any(s <= 0 for s in out_shape):
        raise OracleError("ValueError", "Input cannot be empty")
    if not ax:
        raise OracleError("ValueError", "irfftn needs at least one axis")
    full = reconstruct_hermitian_symmetry(a, out_shape, ax)
    c = ifftn(full, out_shape, ax, norm, engine=engine)  # :712-719
    return c.real.copy()  # :722


# --------------------------------------------------------------------------- strided (strided_fft.rs)


def fft_strided(x, axis: int, engine=None):
    """strided_fft.rs:16-91 (unnormalised forward along one axis)"""
    a = np.asarray(x)
    if axis >= a.ndim or axis < 0:
        raise OracleError("ValueError", f"Axis {axis} is out of bounds for array with {a.ndim} dimensions")
    return _axis_pass(_to_complex(a), axis, False, engine)


fft_strided_complex = fft_strided


def ifft_strided(x, axis: int, engine=None):
    """strided_fft.rs:166-239 (scaled by 1/len(axis))"""
    a = np.asarray(x)
    if axis >= a.ndim or axis < 0:
        raise OracleError("ValueError", f"Axis {axis} is out of bounds for array with {a.ndim} dimensions")
    return _axis_pass(_to_complex(a), axis, True, engine) * (1.0 / a.shape[axis])


# --------------------------------------------------------------------------- backend trait (backend.rs:82-155)


def backend_fft(x, engine=None):
    return _process(_to_complex(np.asarray(x).reshape(-1)), False, engine)


def backend_ifft(x, engine=None):
    a = _to_complex(np.asarray(x).reshape(-1))
    return _process(a, True, engine) * (1.0 / a.size)


def rel_l2(a, b) -> float:
    """relative L2 error ||a - b|| / ||b|| (SURVEY 8d parity metric)"""
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (den if den > 0 else 1.0))
