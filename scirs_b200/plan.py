"""Device-level plan handle — mirror of `FftPlan` / `FftPlanExecutor`
(scirs2-fft/src/planning.rs:75-180, 474-556) with `PlannerBackend::CUDA` (:186) filled in.

`execute_device` takes raw device pointers (or torch CUDA tensors) and a stream:
this is the entry the throughput benchmark times.  `execute` takes host arrays.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from .error import check, ValueError_

_KINDS = {"c2c": _lib.SFC_C2C, "r2c": _lib.SFC_R2C, "c2r": _lib.SFC_C2R}


def _dev_ptr(t):
    if hasattr(t, "data_ptr"):
        return int(t.data_ptr())
    return int(t)


class FftPlan:
    """One cached GPU plan over a C-contiguous N-D array."""

    def __init__(self, shape: Sequence[int], axes: Optional[Sequence[int]] = None, kind: str = "c2c",
                 prec: str = "f64", forward: bool = True, scale: float = 1.0,
                 in_shape: Optional[Sequence[int]] = None, real_input: bool = False, scatter_parts: int = 0,
                 axis_in_len: int = 0, axis_out_len: int = 0, aux_in=None, aux_out=None, real_output: bool = False,
                 dct2: bool = False, dct2_ortho: bool = False, dct3: bool = False, dct4: bool = False,
                 trig_sine: bool = False, scale_dc: float = 0.0):
        lib = _lib.load()
        shape = [int(s) for s in shape]
        if not 1 <= len(shape) <= _lib.SFC_MAX_DIMS:
            raise ValueError_("ndim must be in 1..8")
        axes = list(range(len(shape))) if axes is None else [int(a) for a in axes]
        d = _lib.sfc_desc()
        d.ndim = len(shape)
        for i, s in enumerate(shape):
            d.shape[i] = s
        d.naxes = len(axes)
        for i, a in enumerate(axes):
            d.axes[i] = a
        d.kind = _KINDS[kind]
        d.prec = _lib.SFC_PREC_F64 if prec == "f64" else _lib.SFC_PREC_F32
        d.direction = _lib.SFC_FORWARD if forward else _lib.SFC_INVERSE
        d.scale = float(scale)
        d.flags = 0
        if in_shape is not None:
            d.flags |= _lib.SFC_DESC_CUSTOM_IN_SHAPE
            for i, s in enumerate(in_shape):
                d.in_shape[i] = int(s)
        if real_input:
            d.flags |= _lib.SFC_DESC_REAL_INPUT
        d.scatter_parts = int(scatter_parts)
        if axis_in_len or axis_out_len:  # SFC_DESC_AXIS_LEN: extents of the in / out arrays along the one transformed axis
            d.flags |= _lib.SFC_DESC_AXIS_LEN
            d.axis_in_len, d.axis_out_len = int(axis_in_len), int(axis_out_len)
        if aux_in is not None or aux_out is not None:  # SFC_DESC_AUX_MUL: device tables fused into the load / store
            d.flags |= _lib.SFC_DESC_AUX_MUL
            d.aux_in = None if aux_in is None else _dev_ptr(aux_in)
            d.aux_out = None if aux_out is None else _dev_ptr(aux_out)
            self._keep = (aux_in, aux_out)
        if real_output:
            d.flags |= _lib.SFC_DESC_REAL_OUTPUT
        if dct2:  # kind "r2c" over the last axis: fused DCT-II rows, real in / real out (dct.rs:523-559)
            d.flags |= _lib.SFC_DESC_DCT2 | (_lib.SFC_DESC_DCT2_ORTHO0 if dct2_ortho else 0)
        if dct3:  # fused DCT-III (the inverse packing); scale_dc weighs input 0
            d.flags |= _lib.SFC_DESC_DCT3
        if dct4:  # fused DCT-IV on the n/2-point complex transform (experimental: include/scirs2_fft_cuda.h)
            d.flags |= _lib.SFC_DESC_DCT4
        if trig_sine:  # the sine twins of dct2 / dct3 / dct4
            d.flags |= _lib.SFC_DESC_TRIG_SINE
        d.scale_dc = float(scale_dc)
        self._h = C.c_void_p()
        check(lib.sfc_plan_create(C.byref(self._h), C.byref(d)))
        self._lib = lib
        self.shape, self.axes, self.kind, self.prec, self.forward = tuple(shape), tuple(axes), kind, prec, forward
        info = _lib.sfc_plan_info()
        check(lib.sfc_plan_get_info(self._h, C.byref(info)))
        self.info = {f: getattr(info, f) for f, _ in _lib.sfc_plan_info._fields_}
        # element types of the in / out arrays, from the descriptor flags (not from byte counts)
        cplx = np.complex128 if prec == "f64" else np.complex64
        real = np.float64 if prec == "f64" else np.float32
        real_in = kind == "r2c" or (kind == "c2c" and real_input)
        real_out = kind == "c2r" or (kind == "c2c" and real_output) or (kind == "r2c" and (dct2 or dct3 or dct4))
        self.in_dtype = np.dtype(real if real_in else cplx)
        self.out_dtype = np.dtype(real if real_out else cplx)

    def describe(self) -> str:
        buf = C.create_string_buffer(8192)
        self._lib.sfc_plan_describe(self._h, buf, len(buf))
        return buf.value.decode()

    def execute_device(self, d_in, d_out, stream=0) -> None:
        check(self._lib.sfc_exec_device(self._h, C.c_void_p(_dev_ptr(d_in)), C.c_void_p(_dev_ptr(d_out)),
                                        C.c_void_p(int(stream))))

    def execute(self, x: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Host arrays in/out (H2D + transform + D2H)."""
        a = np.ascontiguousarray(x, dtype=self.in_dtype)
        if a.nbytes != self.info["in_bytes"]:
            raise ValueError_(f"input has {a.nbytes} bytes, plan expects {self.info['in_bytes']}")
        out_dt = self.out_dtype
        if out is None:
            out = np.empty(self.info["out_bytes"] // out_dt.itemsize, dtype=out_dt)
        elif not isinstance(out, np.ndarray) or out.dtype != out_dt or not out.flags.c_contiguous \
                or not out.flags.writeable or out.nbytes != self.info["out_bytes"]:
            # the library writes out_bytes raw bytes: anything else would be overrun or filled with garbage
            raise ValueError_(f"output must be a writable C-contiguous {out_dt} array of {self.info['out_bytes']} bytes")
        check(self._lib.sfc_exec_host(self._h, a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
        return out

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            self._lib.sfc_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FftPlanExecutor:
    """`FftPlanExecutor` — planning.rs:474-556: unnormalised 1-D transform of length prod(shape)."""

    def __init__(self, shape: Sequence[int], forward: bool = True):
        n = 1
        for s in shape:
            n *= int(s)
        self.n = n
        self.plan = FftPlan([n], [0], "c2c", "f64", forward, 1.0)

    def execute(self, input: np.ndarray, output: np.ndarray) -> None:
        if input.size != self.n or output.size != self.n:  # planning.rs:509-517
            raise ValueError_(f"Input size mismatch: expected {self.n}, got {input.size}")
        # `&mut [Complex64]` in the reference: anything that is not a contiguous complex128 buffer cannot be written in place
        if not isinstance(output, np.ndarray) or output.dtype != np.complex128 or not output.flags.c_contiguous:
            raise ValueError_("output must be a C-contiguous complex128 array")
        self.plan.execute(np.asarray(input, dtype=np.complex128), output.reshape(-1))

    def execute_inplace(self, data: np.ndarray) -> None:
        if data.size != self.n:
            raise ValueError_(f"Input size mismatch: expected {self.n}, got {data.size}")
        if not isinstance(data, np.ndarray) or data.dtype != np.complex128 or not data.flags.c_contiguous:
            raise ValueError_("data must be a C-contiguous complex128 array")
        self.plan.execute(np.array(data, dtype=np.complex128), data.reshape(-1))
