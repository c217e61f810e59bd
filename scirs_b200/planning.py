"""Planner objects — mirror of scirs2-fft/src/planning.rs (AdvancedFftPlanner, PlanBuilder,
PlannerBackend, plan_ahead_of_time) and src/planning_parallel.rs (ParallelPlanner,
ParallelExecutor).  Plans are the library's cached GPU plans; `PlannerBackend.CUDA` (planning.rs:186)
is the only backend this package provides.
"""
from __future__ import annotations

import ctypes as C
import enum
import time
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from .error import check, ValueError_
from .plan import FftPlan, FftPlanExecutor


class PlannerBackend(enum.Enum):
    """planning.rs:179-190"""
    RustFFT = "rustfft"
    FFTW = "fftw"
    CUDA = "cuda"
    Custom = "custom"


class PlanningStrategy(enum.Enum):
    """planning.rs:28-40"""
    AlwaysNew = 0
    CacheFirst = 1
    SerializedFirst = 2
    AutoTuned = 3


class AdvancedFftPlanner:
    """planning.rs:203-456: plans keyed (shape, forward, backend); 1-D transform of prod(shape)."""

    def __init__(self):
        self._plans = {}

    def plan_fft(self, shape: Sequence[int], forward: bool = True,
                 backend: PlannerBackend = PlannerBackend.CUDA) -> FftPlanExecutor:
        if backend not in (PlannerBackend.CUDA, PlannerBackend.RustFFT):
            from .error import NotImplementedError_

            raise NotImplementedError_(f"backend {backend.name} is not available")
        key = (tuple(int(s) for s in shape), bool(forward))
        ex = self._plans.get(key)
        if ex is None:
            ex = FftPlanExecutor(shape, forward)
            self._plans[key] = ex
        return ex

    def clear_cache(self) -> None:
        self._plans.clear()


_PLANNER = AdvancedFftPlanner()


def get_global_planner() -> AdvancedFftPlanner:
    return _PLANNER


def plan_ahead_of_time(sizes: Sequence[int], db_path: Optional[str] = None) -> None:
    """planning.rs:671-693: create forward plans for the listed sizes up front."""
    for n in sizes:
        get_global_planner().plan_fft([int(n)], True)


class PlanBuilder:
    """planning.rs:560-668"""

    def __init__(self):
        self._shape = None
        self._forward = True
        self._backend = PlannerBackend.CUDA

    def shape(self, shape: Sequence[int]) -> "PlanBuilder":
        self._shape = [int(s) for s in shape]
        return self

    def forward(self, forward: bool) -> "PlanBuilder":
        self._forward = bool(forward)
        return self

    def backend(self, backend: PlannerBackend) -> "PlanBuilder":
        self._backend = backend
        return self

    def build(self) -> FftPlanExecutor:
        if self._shape is None:
            raise ValueError_("Shape must be specified")
        return get_global_planner().plan_fft(self._shape, self._forward, self._backend)


class ParallelExecutor:
    """planning_parallel.rs:246-405: one plan over many equally sized signals."""

    def __init__(self, shape: Sequence[int], forward: bool = True):
        self.size = 1
        for s in shape:
            self.size *= int(s)
        self.forward = bool(forward)
        self._single = FftPlanExecutor([self.size], forward)

    def execute(self, input: np.ndarray, output: np.ndarray) -> None:
        self._single.execute(np.asarray(input), output)

    def execute_batch(self, inputs: Sequence[np.ndarray], outputs: List[np.ndarray]) -> List[float]:
        if len(inputs) != len(outputs):  # planning_parallel.rs:321-325
            raise ValueError_("Input and output counts must match")
        for i, (a, b) in enumerate(zip(inputs, outputs)):
            if a.size != self.size:  # :330-346
                raise ValueError_(f"Input {i} has wrong size: expected {self.size}, got {a.size}")
            if b.size != self.size:
                raise ValueError_(f"Output {i} has wrong size: expected {self.size}, got {b.size}")
        if not inputs:
            return []
        t0 = time.perf_counter()
        stacked = np.ascontiguousarray(np.stack([np.asarray(a, dtype=np.complex128).reshape(-1) for a in inputs]))
        res = np.empty_like(stacked)
        check(_lib.load().sfc_execute_batch(stacked.ctypes.data_as(C.c_void_p), res.ctypes.data_as(C.c_void_p),
                                            len(inputs), self.size, 0 if self.forward else 1))
        for o, r in zip(outputs, res):
            o.reshape(-1)[:] = r
        dt = (time.perf_counter() - t0) / len(inputs)
        return [dt] * len(inputs)  # the reference returns one Duration per item


class ParallelPlanner:
    """planning_parallel.rs:75-244"""

    def __init__(self):
        self._planner = AdvancedFftPlanner()

    def plan_fft(self, shape: Sequence[int], forward: bool = True,
                 backend: PlannerBackend = PlannerBackend.CUDA) -> FftPlanExecutor:
        return self._planner.plan_fft(shape, forward, backend)

    def plan_multiple(self, specs: Sequence[tuple]) -> List[FftPlanExecutor]:
        return [self.plan_fft(*spec) for spec in specs]

    def clear_cache(self) -> None:
        self._planner.clear_cache()


# ------------------------------------------------------------------ planning_adaptive.rs


class AdaptivePlanningConfig:
    """planning_adaptive.rs:13-45"""

    def __init__(self, enabled: bool = True, min_samples: int = 5, evaluation_interval: float = 10.0,
                 max_strategy_switches: int = 3, enable_backend_switching: bool = True, improvement_threshold: float = 1.1):
        self.enabled = enabled
        self.min_samples = min_samples
        self.evaluation_interval = evaluation_interval  # seconds (a Duration in the reference)
        self.max_strategy_switches = max_strategy_switches
        self.enable_backend_switching = enable_backend_switching
        self.improvement_threshold = improvement_threshold


class _StrategyMetrics:
    """planning_adaptive.rs:48-86: total / count / integer-nanosecond average"""

    def __init__(self):
        self.total_ns, self.count, self.avg_ns = 0, 0, 0

    def record(self, seconds: float) -> None:
        self.total_ns += int(round(seconds * 1e9))
        self.count += 1
        self.avg_ns = self.total_ns // self.count


class AdaptivePlanner:
    """planning_adaptive.rs:86-252: records execution times per `PlanningStrategy` and switches to a strategy whose
    average beats the current one by `improvement_threshold`, at most `max_strategy_switches` times.  The plan itself is
    the library's cached GPU plan whatever the strategy (the strategies differ in where the HOST looks for a plan first)."""

    def __init__(self, size: Sequence[int], forward: bool = True, config: Optional[AdaptivePlanningConfig] = None):
        self.size = [int(s) for s in size]
        self.forward = bool(forward)
        self.config = config or AdaptivePlanningConfig()
        self._strategy = PlanningStrategy.CacheFirst  # :131 "Start with a reasonable default"
        self._backend = PlannerBackend.CUDA
        self._metrics = {s: _StrategyMetrics() for s in PlanningStrategy}
        self._last_switch = time.monotonic()
        self._switches = 0
        self._plan: Optional[FftPlanExecutor] = None

    def current_strategy(self) -> PlanningStrategy:
        return self._strategy

    def current_backend(self) -> PlannerBackend:
        return self._backend

    def get_plan(self) -> FftPlanExecutor:
        """:153-172"""
        if self._plan is None:
            self._plan = AdvancedFftPlanner().plan_fft(self.size, self.forward, self._backend)
        return self._plan

    def record_execution(self, execution_time: float) -> None:
        """:175-199 (seconds)"""
        if not self.config.enabled:
            return
        m = self._metrics[self._strategy]
        m.record(execution_time)
        if (m.count >= self.config.min_samples and time.monotonic() - self._last_switch >= self.config.evaluation_interval
                and self._switches < self.config.max_strategy_switches):
            self._evaluate_strategies()

    def _evaluate_strategies(self) -> None:
        """:202-238"""
        best, best_ns = self._strategy, self._metrics[self._strategy].avg_ns
        for strat, m in self._metrics.items():
            if m.count == 0 or m.avg_ns == 0:
                continue
            if best_ns / m.avg_ns > self.config.improvement_threshold:
                best, best_ns = strat, m.avg_ns
        if best != self._strategy:
            self._strategy = best
            self._last_switch = time.monotonic()
            self._switches += 1
            self._plan = None

    def get_statistics(self):
        """:241-250: {strategy: (average seconds, count)}"""
        return {s: (m.avg_ns * 1e-9, m.count) for s, m in self._metrics.items()}


class AdaptiveExecutor:
    """planning_adaptive.rs:254-316"""

    def __init__(self, size: Sequence[int], forward: bool = True, config: Optional[AdaptivePlanningConfig] = None):
        self._planner = AdaptivePlanner(size, forward, config)

    def execute(self, input: np.ndarray, output: np.ndarray) -> None:
        t0 = time.perf_counter()
        self._planner.get_plan().execute(input, output)
        self._planner.record_execution(time.perf_counter() - t0)

    def current_strategy(self) -> PlanningStrategy:
        return self._planner.current_strategy()

    def get_statistics(self):
        return self._planner.get_statistics()
