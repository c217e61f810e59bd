"""`FftBackend` trait, the CUDA backend and `BackendManager` — mirror of
scirs2-fft/src/backend.rs:14-48, 163-344 (the drop-in boundary).
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Dict

import numpy as np

from . import _lib
from .error import check, ValueError_


class FftBackend:
    """backend.rs:14-48"""

    def name(self) -> str:
        raise NotImplementedError

    def description(self) -> str:
        raise NotImplementedError

    def is_available(self) -> bool:
        raise NotImplementedError

    def fft(self, input: np.ndarray, output: np.ndarray) -> None:
        raise NotImplementedError

    def ifft(self, input: np.ndarray, output: np.ndarray) -> None:
        raise NotImplementedError

    def fft_sized(self, input: np.ndarray, output: np.ndarray, size: int) -> None:
        raise NotImplementedError

    def ifft_sized(self, input: np.ndarray, output: np.ndarray, size: int) -> None:
        raise NotImplementedError

    def supports_feature(self, feature: str) -> bool:
        raise NotImplementedError


def _c128(a: np.ndarray, writable=False) -> np.ndarray:
    if a.dtype != np.complex128 or not a.flags.c_contiguous:
        if writable:
            raise ValueError_("output must be a contiguous complex128 array")
        a = np.ascontiguousarray(a, dtype=np.complex128)
    return a


class CudaFftBackend(FftBackend):
    """The backend a maintainer registers as "cuda_fft" (examples/backend_example.rs:103-104)."""

    def name(self) -> str:
        return _lib.load().sfc_backend_name().decode()

    def description(self) -> str:
        return _lib.load().sfc_backend_description().decode()

    def is_available(self) -> bool:
        return bool(_lib.load().sfc_is_available())

    def fft(self, input, output) -> None:
        i, o = _c128(np.asarray(input)), _c128(output, True)
        check(_lib.load().sfc_backend_fft(i.ctypes.data_as(C.c_void_p), i.size, o.ctypes.data_as(C.c_void_p), o.size))

    def ifft(self, input, output) -> None:
        i, o = _c128(np.asarray(input)), _c128(output, True)
        check(_lib.load().sfc_backend_ifft(i.ctypes.data_as(C.c_void_p), i.size, o.ctypes.data_as(C.c_void_p), o.size))

    def fft_sized(self, input, output, size: int) -> None:
        i, o = _c128(np.asarray(input)), _c128(output, True)
        check(_lib.load().sfc_backend_fft_sized(i.ctypes.data_as(C.c_void_p), i.size, o.ctypes.data_as(C.c_void_p),
                                                o.size, int(size)))

    def ifft_sized(self, input, output, size: int) -> None:
        i, o = _c128(np.asarray(input)), _c128(output, True)
        check(_lib.load().sfc_backend_ifft_sized(i.ctypes.data_as(C.c_void_p), i.size, o.ctypes.data_as(C.c_void_p),
                                                 o.size, int(size)))

    def supports_feature(self, feature: str) -> bool:
        return bool(_lib.load().sfc_backend_supports_feature(feature.encode()))


class BackendManager:
    """backend.rs:163-281 — registry with one current backend."""

    def __init__(self):
        self._lock = threading.Lock()
        self._backends: Dict[str, FftBackend] = {"cuda_fft": CudaFftBackend()}
        self._current = "cuda_fft"

    def list_backends(self):
        with self._lock:
            return list(self._backends)

    def get_backend_name(self) -> str:
        with self._lock:
            return self._current

    def register_backend(self, name: str, backend: FftBackend) -> None:
        with self._lock:
            if name in self._backends:  # backend.rs:184-194
                raise ValueError_(f"Backend '{name}' already exists")
            self._backends[name] = backend

    def set_backend(self, name: str) -> None:
        with self._lock:
            b = self._backends.get(name)
            if b is None:  # backend.rs:203-224
                raise ValueError_(f"Backend '{name}' not found")
            if not b.is_available():
                raise ValueError_(f"Backend '{name}' is not available")
            self._current = name

    def get_backend(self) -> FftBackend:
        with self._lock:
            return self._backends[self._current]

    def get_backend_info(self, name: str):
        with self._lock:
            b = self._backends.get(name)
        if b is None:
            return None
        return {"name": b.name(), "description": b.description(), "available": b.is_available()}


_MANAGER = BackendManager()


def get_backend_manager() -> BackendManager:
    return _MANAGER


class BackendContext:
    """RAII guard of backend.rs:318-344 as a context manager."""

    def __init__(self, name: str):
        self._name = name
        self._prev = None

    def __enter__(self):
        m = get_backend_manager()
        self._prev = m.get_backend_name()
        m.set_backend(self._name)
        return self

    def __exit__(self, *exc):
        get_backend_manager().set_backend(self._prev)
        return False
