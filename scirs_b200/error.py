"""FFTError hierarchy — mirror of scirs2-fft/src/error.rs:7-46.

Every negative ``sfc_status`` returned across the C ABI is raised as the
matching variant with the library's message text.
"""
from __future__ import annotations

from . import _lib


class FFTError(Exception):
    """Base of all errors (the reference's ``FFTError`` enum)."""


class ComputationError(FFTError):
    pass


class DimensionError(FFTError):
    pass


class ValueError_(FFTError, ValueError):
    """``FFTError::ValueError`` (also a Python ValueError)."""


class NotImplementedError_(FFTError, NotImplementedError):
    """``FFTError::NotImplementedError``."""


class IOError_(FFTError):
    pass


class BackendError(FFTError):
    pass


class PlanError(FFTError):
    pass


class CommunicationError(FFTError):
    pass


class MemoryError_(FFTError):
    pass


_BY_CODE = {
    _lib.SFC_ERR_COMPUTATION: ComputationError,
    _lib.SFC_ERR_DIMENSION: DimensionError,
    _lib.SFC_ERR_VALUE: ValueError_,
    _lib.SFC_ERR_NOT_IMPLEMENTED: NotImplementedError_,
    _lib.SFC_ERR_IO: IOError_,
    _lib.SFC_ERR_BACKEND: BackendError,
    _lib.SFC_ERR_PLAN: PlanError,
    _lib.SFC_ERR_COMMUNICATION: CommunicationError,
    _lib.SFC_ERR_MEMORY: MemoryError_,
}


def check(rc: int) -> None:
    """Raise the FFTError variant for a negative status code."""
    if rc >= 0:
        return
    lib = _lib.load()
    msg = lib.sfc_last_error().decode("utf-8", "replace")
    raise _BY_CODE.get(rc, FFTError)(msg)
