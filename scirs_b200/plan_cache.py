"""Plan cache controls — mirror of scirs2-fft/src/plan_cache.rs:28-235.

The cache itself lives inside libscirs2_fft_cuda.so (keyed by the full plan
descriptor + device, 128 entries, 1 h TTL, LRU); this is its handle.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

from . import _lib
from .error import check


@dataclass
class CacheStats:
    """plan_cache.rs:199-220"""
    hit_count: int
    miss_count: int
    hit_rate: float
    size: int
    max_size: int


class PlanCache:
    def set_enabled(self, enabled: bool) -> None:
        check(_lib.load().sfc_cache_set_enabled(1 if enabled else 0))

    def is_enabled(self) -> bool:
        return bool(_lib.load().sfc_cache_is_enabled())

    def clear(self) -> None:
        check(_lib.load().sfc_cache_clear())

    def configure(self, max_entries: int = 128, max_age_seconds: float = 3600.0) -> None:
        """`PlanCache::with_config` (plan_cache.rs:47-54): fresh cache with new limits."""
        check(_lib.load().sfc_cache_configure(int(max_entries), float(max_age_seconds)))

    def get_stats(self) -> CacheStats:
        s = _lib.sfc_cache_stats()
        check(_lib.load().sfc_cache_get_stats(C.byref(s)))
        return CacheStats(s.hit_count, s.miss_count, s.hit_rate, s.size, s.max_size)


_GLOBAL = PlanCache()


def get_global_cache() -> PlanCache:
    """plan_cache.rs:223-228"""
    return _GLOBAL
