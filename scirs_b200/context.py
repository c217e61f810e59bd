"""Workers / context configuration holders — mirror of scirs2-fft/src/worker_pool.rs and
src/context.rs.  As in the reference, `workers` is advisory: the reference's
`execute*` run inline and `set_workers` is a no-op (worker_pool.rs:102-122, 208-213); on the GPU the
unit of parallelism is the device, so the value is recorded and otherwise unused.
"""
from __future__ import annotations

import os
import threading
from contextlib import contextmanager
from dataclasses import dataclass, field
from typing import Callable, Optional

from .backend import get_backend_manager
from .plan_cache import get_global_cache


def _default_workers() -> int:
    """worker_pool.rs:27-35: available parallelism unless SCIRS2_FFT_WORKERS overrides it."""
    env = os.environ.get("SCIRS2_FFT_WORKERS")
    if env is not None:
        try:
            return int(env)
        except ValueError:
            pass
    return os.cpu_count() or 1


@dataclass
class WorkerConfig:
    """worker_pool.rs:12-43"""
    num_workers: int = field(default_factory=_default_workers)
    enabled: bool = True
    stack_size: Optional[int] = None
    thread_name_prefix: str = "scirs2-fft-worker"


@dataclass
class WorkerPoolInfo:
    """worker_pool.rs:143-159"""
    num_workers: int
    enabled: bool
    thread_name_prefix: str


class WorkerPool:
    """worker_pool.rs:46-141"""

    def __init__(self, config: Optional[WorkerConfig] = None):
        self._lock = threading.Lock()
        self._config = config or WorkerConfig()

    def get_workers(self) -> int:
        with self._lock:
            return self._config.num_workers

    def set_workers(self, num_workers: int) -> None:
        with self._lock:
            self._config.num_workers = int(num_workers)

    def is_enabled(self) -> bool:
        with self._lock:
            return self._config.enabled

    def set_enabled(self, enabled: bool) -> None:
        with self._lock:
            self._config.enabled = bool(enabled)

    def execute(self, f: Callable):
        return f()  # worker_pool.rs:102-112 runs inline

    def execute_with_workers(self, _num_workers: int, f: Callable):
        return f()  # worker_pool.rs:114-122

    def get_info(self) -> WorkerPoolInfo:
        with self._lock:
            return WorkerPoolInfo(self._config.num_workers, self._config.enabled, self._config.thread_name_prefix)


_POOL = WorkerPool()


def get_global_pool() -> WorkerPool:
    return _POOL


def set_workers(_n: int) -> None:
    """worker_pool.rs:208-213: accepted, no effect on the global pool (it is immutable once created)."""
    return None


def get_workers() -> int:
    return get_global_pool().get_workers()


class FftContext:
    """context.rs:13-84: scoped backend / workers / cache settings, restored on exit."""

    def __init__(self):
        self._backend: Optional[str] = None
        self._workers: Optional[int] = None
        self._cache: Optional[bool] = None
        self._prev_backend = None
        self._prev_cache = None

    def with_backend(self, name: str) -> "FftContext":
        get_backend_manager().get_backend_info(name)
        if name not in get_backend_manager().list_backends():
            from .error import ValueError_

            raise ValueError_(f"Backend '{name}' not found")
        self._backend = name
        return self

    def with_workers(self, num_workers: int) -> "FftContext":
        self._workers = int(num_workers)
        return self

    def with_cache(self, enabled: bool) -> "FftContext":
        self._cache = bool(enabled)
        return self

    def __enter__(self) -> "FftContext":
        if self._backend is not None:
            self._prev_backend = get_backend_manager().get_backend_name()
            get_backend_manager().set_backend(self._backend)
        if self._cache is not None:
            self._prev_cache = get_global_cache().is_enabled()
            get_global_cache().set_enabled(self._cache)
        return self

    def __exit__(self, *exc) -> bool:
        if self._prev_backend is not None:
            get_backend_manager().set_backend(self._prev_backend)
        if self._prev_cache is not None:
            get_global_cache().set_enabled(self._prev_cache)
        return False


class FftContextBuilder:
    """context.rs:87-158"""

    def __init__(self):
        self._backend = None
        self._workers = None
        self._cache_enabled = None
        self._cache_size = None
        self._cache_ttl = None

    def backend(self, name: str) -> "FftContextBuilder":
        self._backend = name
        return self

    def workers(self, count: int) -> "FftContextBuilder":
        self._workers = count
        return self

    def cache_enabled(self, enabled: bool) -> "FftContextBuilder":
        self._cache_enabled = enabled
        return self

    def cache_size(self, size: int) -> "FftContextBuilder":
        self._cache_size = size
        return self

    def cache_ttl(self, seconds: float) -> "FftContextBuilder":
        self._cache_ttl = seconds
        return self

    def build(self) -> FftContext:
        ctx = FftContext()
        if self._backend is not None:
            ctx.with_backend(self._backend)
        if self._workers is not None:
            ctx.with_workers(self._workers)
        if self._cache_enabled is not None:
            ctx.with_cache(self._cache_enabled)
        if self._cache_size is not None or self._cache_ttl is not None:
            get_global_cache().configure(self._cache_size or 128, self._cache_ttl or 3600.0)
        return ctx


def fft_context() -> FftContextBuilder:
    return FftContextBuilder()


def with_fft_settings(builder: FftContextBuilder, f: Callable):
    with builder.build():
        return f()


def with_backend(backend: str, f: Callable):
    """context.rs:193-198"""
    with FftContext().with_backend(backend):
        return f()


def with_workers(workers: int, f: Callable):
    """context.rs:201-206 / worker_pool.rs:221-227"""
    with FftContext().with_workers(workers):
        return f()


def without_cache(f: Callable):
    """context.rs:209-214"""
    with FftContext().with_cache(False):
        return f()
