"""Auto-tuning mirror — scirs2-fft/src/auto_tuning.rs:25-600.

Same types, defaults and selection rules as the reference (`SizeRange`, `SizeStep`, `FftVariant`,
`AutoTuneConfig`, `BenchmarkResult`, `SystemInfo`, `TuningDatabase`, `AutoTuner`); the timed work is the GPU
plan of this library instead of a rustfft plan.  What the variants mean here:

* ``Standard``   — plan looked up / built with the plan cache DISABLED for the call (planning cost included,
                   as `FftPlanner::new()` per call is in the reference, auto_tuning.rs:349-360);
* ``Cached``     — plan taken from the process-wide plan cache (the reference's `create_and_time_plan` path);
* ``InPlace``    — cached plan, result written over the input buffer (`process_with_scratch`, :372-381);
* ``SplitRadix`` — placeholder in the reference (:392-404: falls back to the standard plan); same here.

`best_algorithms` is keyed by `(size, forward)`; JSON has no tuple keys, so it is stored as a list of
`[[size, forward], variant]` pairs.
"""
from __future__ import annotations

import enum
import json
import math
import os
import platform
import time
from dataclasses import asdict, dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .error import IOError_, ValueError_


class FftVariant(str, enum.Enum):
    """auto_tuning.rs:49-58"""
    Standard = "Standard"
    InPlace = "InPlace"
    Cached = "Cached"
    SplitRadix = "SplitRadix"


@dataclass
class SizeStep:
    """auto_tuning.rs:36-45: kind in {"Linear", "Exponential", "PowersOfTwo", "Custom"}"""
    kind: str = "PowersOfTwo"
    value: object = None

    @staticmethod
    def Linear(step: int) -> "SizeStep":
        return SizeStep("Linear", int(step))

    @staticmethod
    def Exponential(factor: float) -> "SizeStep":
        return SizeStep("Exponential", float(factor))

    @staticmethod
    def PowersOfTwo() -> "SizeStep":
        return SizeStep("PowersOfTwo")

    @staticmethod
    def Custom(sizes: Sequence[int]) -> "SizeStep":
        return SizeStep("Custom", [int(s) for s in sizes])


@dataclass
class SizeRange:
    """auto_tuning.rs:25-32"""
    min: int = 16
    max: int = 8192
    step: SizeStep = field(default_factory=SizeStep.PowersOfTwo)

    def sizes(self) -> List[int]:
        """auto_tuning.rs:296-338 `generate_sizes`"""
        k, v = self.step.kind, self.step.value
        out: List[int] = []
        if k == "Linear":
            s = self.min
            while s <= self.max:
                out.append(s)
                s += max(int(v), 1)
        elif k == "Exponential":
            s = float(self.min)
            while s <= self.max:
                out.append(int(s))
                s *= max(float(v), 1.0 + 1e-9)
        elif k == "PowersOfTwo":
            s = 1
            while s < self.min:
                s *= 2
            while s <= self.max:
                out.append(s)
                s *= 2
        elif k == "Custom":
            out = [s for s in v if self.min <= s <= self.max]
        else:
            raise ValueError_(f"unknown SizeStep {k}")
        return out


@dataclass
class AutoTuneConfig:
    """auto_tuning.rs:62-89 (defaults: 16..8192 powers of two, 10 repetitions, 3 warm-ups, Standard + Cached)"""
    sizes: SizeRange = field(default_factory=SizeRange)
    repetitions: int = 10
    warmup: int = 3
    variants: List[FftVariant] = field(default_factory=lambda: [FftVariant.Standard, FftVariant.Cached])
    database_path: str = ".fft_tuning_db.json"


@dataclass
class SystemInfo:
    """auto_tuning.rs:112-121, with the GPU target appended to the feature list"""
    cpu_model: str = ""
    num_cores: int = 0
    architecture: str = ""
    cpu_features: List[str] = field(default_factory=list)

    @staticmethod
    def detect() -> "SystemInfo":
        return SystemInfo(platform.processor() or "unknown", os.cpu_count() or 1, platform.machine(), ["sm_100a"])


@dataclass
class BenchmarkResult:
    """auto_tuning.rs:93-108"""
    size: int
    variant: FftVariant
    forward: bool
    avg_time_ns: int
    min_time_ns: int
    std_dev_ns: float
    system_info: SystemInfo


@dataclass
class TuningDatabase:
    """auto_tuning.rs:125-131"""
    results: List[BenchmarkResult] = field(default_factory=list)
    last_updated: int = 0
    best_algorithms: Dict[Tuple[int, bool], FftVariant] = field(default_factory=dict)

    def to_json(self) -> dict:
        return {"results": [dict(asdict(r), variant=r.variant.value) for r in self.results], "last_updated": self.last_updated,
                "best_algorithms": [[[s, f], v.value] for (s, f), v in self.best_algorithms.items()]}

    @staticmethod
    def from_json(d: dict) -> "TuningDatabase":
        db = TuningDatabase(last_updated=int(d.get("last_updated", 0)))
        for r in d.get("results", []):
            db.results.append(BenchmarkResult(int(r["size"]), FftVariant(r["variant"]), bool(r["forward"]), int(r["avg_time_ns"]),
                                              int(r["min_time_ns"]), float(r["std_dev_ns"]), SystemInfo(**r["system_info"])))
        for (s, f), v in d.get("best_algorithms", []):
            db.best_algorithms[(int(s), bool(f))] = FftVariant(v)
        return db


class AutoTuner:
    """auto_tuning.rs:135-600"""

    def __init__(self, config: Optional[AutoTuneConfig] = None):
        self.config = config or AutoTuneConfig()
        self.enabled = True
        try:
            self.database = self._load(self.config.database_path)
        except Exception:  # :158-166: any failure = empty database
            self.database = TuningDatabase(last_updated=int(time.time()))

    @staticmethod
    def with_config(config: AutoTuneConfig) -> "AutoTuner":
        return AutoTuner(config)

    @staticmethod
    def _load(path: str) -> TuningDatabase:
        if not os.path.exists(path):
            raise IOError_("Tuning database file does not exist")
        with open(path) as f:
            return TuningDatabase.from_json(json.load(f))

    def save_database(self) -> None:
        """:195-216"""
        parent = os.path.dirname(self.config.database_path)
        try:
            if parent:
                os.makedirs(parent, exist_ok=True)
            with open(self.config.database_path, "w") as f:
                json.dump(self.database.to_json(), f, indent=2)
        except OSError as e:
            raise IOError_(f"Failed to create tuning database file: {e}")

    def set_enabled(self, enabled: bool) -> None:
        self.enabled = bool(enabled)

    def is_enabled(self) -> bool:
        return self.enabled

    # -------------------------------------------------------------- benchmarking (needs a CUDA device)
    def _run_variant(self, buf: np.ndarray, variant: FftVariant, forward: bool) -> np.ndarray:
        from . import plan_cache
        from .plan import FftPlan

        n = buf.size
        if variant in (FftVariant.Standard, FftVariant.SplitRadix):
            cache = plan_cache.get_global_cache()
            was = cache.is_enabled()
            cache.set_enabled(False)
            try:
                return FftPlan([n], [0], "c2c", "f64", forward).execute(buf)
            finally:
                cache.set_enabled(was)
        plan = FftPlan([n], [0], "c2c", "f64", forward)
        if variant == FftVariant.InPlace:
            return plan.execute(buf, buf)
        return plan.execute(buf)

    def benchmark_variant(self, size: int, variant: FftVariant, forward: bool) -> BenchmarkResult:
        """:253-293: warm-ups, then `repetitions` timed runs of one transform"""
        rng = np.random.default_rng(size)
        times = []
        for it in range(self.config.warmup + self.config.repetitions):
            buf = (rng.standard_normal(size) + 1j * rng.standard_normal(size)).astype(np.complex128)
            t0 = time.perf_counter_ns()
            self._run_variant(buf, variant, forward)
            dt = time.perf_counter_ns() - t0
            if it >= self.config.warmup:
                times.append(dt)
        avg = sum(times) / len(times)
        var = sum((t - avg) ** 2 for t in times) / len(times)
        return BenchmarkResult(size, variant, forward, int(avg), int(min(times)), math.sqrt(var), SystemInfo.detect())

    def run_benchmarks(self) -> None:
        """:229-250"""
        if not self.enabled:
            return
        for size in self.config.sizes.sizes():
            for variant in self.config.variants:
                for forward in (True, False):
                    self.database.results.append(self.benchmark_variant(size, variant, forward))
        self.update_best_algorithms()
        self.database.last_updated = int(time.time())
        self.save_database()

    def update_best_algorithms(self) -> None:
        """:441-475: per (size, forward) the variant with the smallest average time"""
        self.database.best_algorithms.clear()
        best: Dict[Tuple[int, bool], BenchmarkResult] = {}
        for r in self.database.results:
            k = (r.size, r.forward)
            if k not in best or r.avg_time_ns < best[k].avg_time_ns:
                best[k] = r
        for k, r in best.items():
            self.database.best_algorithms[k] = r.variant

    def get_best_variant(self, size: int, forward: bool) -> FftVariant:
        """:478-510: exact size, else the closest tuned size of the same direction, else Standard"""
        if not self.enabled:
            return FftVariant.Standard
        ba = self.database.best_algorithms
        if (size, forward) in ba:
            return ba[(size, forward)]
        closest, min_diff = 0, None
        for (s, f) in ba:
            if f == forward:
                d = abs(s - size)
                if min_diff is None or d < min_diff:
                    closest, min_diff = s, d
        if closest > 0:
            return ba.get((closest, forward), FftVariant.Standard)
        return FftVariant.Standard

    def run_optimal_fft(self, input, size: Optional[int] = None, forward: bool = True) -> np.ndarray:
        """:513-600: zero-pad to `size`, transform with the best variant; UNNORMALISED in both directions
        (the reference calls `Fft::process` directly)."""
        a = np.asarray(input, dtype=np.complex128).reshape(-1)
        n = a.size if size is None else int(size)
        buf = np.zeros(max(n, a.size), dtype=np.complex128)
        buf[: a.size] = a
        buf = buf[:n] if n <= a.size else buf  # the reference only ever pads (buffer.len() < actual_size)
        return np.asarray(self._run_variant(np.ascontiguousarray(buf), self.get_best_variant(n, forward), forward))
