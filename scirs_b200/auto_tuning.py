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


# ------------------------------------------------------------------ GPU planner tuning (round 2)


class GpuPlanTuner:
    """Times the GPU planner's variants of ONE plan on the device and persists the winner — the reference's
    `AutoTuner::run_benchmarks` / `TuningDatabase` idea (auto_tuning.rs:188-229, 409-470) applied to what actually varies here:
    tile shapes, pass counts and the TMA flavours (the `SFC_*` options of DESIGN.md section 11), not CPU algorithm variants.

    Options are switched at run time through `sfc_planner_set_option` (no new process per variant).  The database is JSON
    next to the reference's own (`database_path`), keyed by `arch|kind|prec|shape|axes|direction`; `plan()` builds a plan
    under the stored options, so `plan_ahead_of_time`-style warm-up gets the tuned tiles.
    """

    def __init__(self, database_path: str = ".fft_gpu_tuning_db.json", repetitions: int = 20, warmup: int = 3):
        self.database_path = database_path
        self.repetitions, self.warmup = int(repetitions), int(warmup)
        self.entries: Dict[str, dict] = {}
        self.arch_id = f"{platform.machine()}-sm_100a"
        if os.path.exists(database_path):
            self.load()

    # -- persistence
    def load(self) -> None:
        try:
            d = json.load(open(self.database_path))
        except (OSError, ValueError) as ex:
            raise IOError_(f"cannot read {self.database_path}: {ex}")
        if d.get("arch_id") == self.arch_id:  # winners of another GPU generation do not transfer
            self.entries = dict(d.get("entries", {}))

    def save(self) -> None:
        try:
            json.dump({"arch_id": self.arch_id, "entries": self.entries}, open(self.database_path, "w"), indent=1, sort_keys=True)
        except OSError as ex:
            raise IOError_(f"cannot write {self.database_path}: {ex}")

    # -- candidates
    @staticmethod
    def key(shape, axes, kind, prec, forward) -> str:
        return "|".join(["sm_100a", kind, prec, "x".join(str(int(s)) for s in shape), ",".join(str(int(a)) for a in axes),
                         "fwd" if forward else "inv"])

    @staticmethod
    def candidates(shape: Sequence[int], axes: Sequence[int]) -> List[Dict[str, str]]:
        """Option sets worth timing for this geometry ({} = the library defaults, always first)."""
        cands: List[Dict[str, str]] = [{}]
        last = len(shape) - 1
        for a in axes:
            n = int(shape[a])
            pow2 = n & (n - 1) == 0
            if pow2 and a == last and n >= 2048:          # contiguous rows: persistent late-prefetch flavour
                cands += [{"SFC_PIPE_LATE": "2"}]
            if pow2 and a == last and n > 8192:           # multi-pass rows: two passes of big tiles / three of small ones, TMA on or off
                cands += [{"SFC_PIPE_LATE": "0"}, {"SFC_PIPE_LATE": "3"}, {"SFC_THREE_LEVEL_MIN": str(n)},
                          {"SFC_THREE_LEVEL_MIN": str(1 << 40)}]
            if pow2 and a != last:                        # strided axis: tensor-map tiles, narrower tiles
                cands += [{"SFC_PIPE_LATE": "3"}, {"SFC_COL_SMEM_KB": "40"}]
            if not pow2 and n > 4096:                     # Bluestein: first factor of the padded length
                cands += [{"SFC_BLUE_L1": v} for v in ("256", "512", "2048")]
        out, seen = [], set()
        for c in cands:
            t = tuple(sorted(c.items()))
            if t not in seen:
                seen.add(t)
                out.append(c)
        return out

    # -- measurement
    def _time(self, shape, axes, kind, prec, forward, options: Dict[str, str]) -> float:
        import torch

        from . import _lib
        from .error import check
        from .plan import FftPlan

        lib = _lib.load()
        for k, v in options.items():
            check(lib.sfc_planner_set_option(k.encode(), v.encode()))
        try:
            plan = FftPlan(shape, axes, kind, prec, forward)
            rt = torch.float64 if prec == "f64" else torch.float32
            x = torch.randn(plan.info["in_bytes"] // (8 if prec == "f64" else 4), dtype=rt, device="cuda")
            y = torch.empty(plan.info["out_bytes"] // (8 if prec == "f64" else 4), dtype=rt, device="cuda")
            st = torch.cuda.current_stream()
            for _ in range(self.warmup):
                plan.execute_device(x, y, st.cuda_stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(self.repetitions):
                plan.execute_device(x, y, st.cuda_stream)
            e1.record(st)
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / self.repetitions
        finally:
            for k in options:
                lib.sfc_planner_set_option(k.encode(), None)

    def tune(self, shape: Sequence[int], axes: Optional[Sequence[int]] = None, kind: str = "c2c", prec: str = "f64",
             forward: bool = True) -> dict:
        shape = [int(s) for s in shape]
        axes = list(range(len(shape))) if axes is None else [int(a) for a in axes]
        results = []
        for opts in self.candidates(shape, axes):
            try:
                results.append((self._time(shape, axes, kind, prec, forward, opts), opts))
            except Exception as ex:  # an option set the planner refuses for this shape is simply not a candidate
                results.append((float("inf"), dict(opts, error=str(ex)[:80])))
        best_ms, best = min(results, key=lambda r: r[0])
        entry = {"options": best, "ms": round(best_ms, 5), "default_ms": round(results[0][0], 5),
                 "candidates": [{"options": o, "ms": (None if math.isinf(t) else round(t, 5))} for t, o in results]}
        self.entries[self.key(shape, axes, kind, prec, forward)] = entry
        return entry

    def options_for(self, shape, axes=None, kind="c2c", prec="f64", forward=True) -> Dict[str, str]:
        axes = list(range(len(shape))) if axes is None else list(axes)
        e = self.entries.get(self.key(shape, axes, kind, prec, forward))
        return dict(e["options"]) if e else {}

    def plan(self, shape, axes=None, kind="c2c", prec="f64", forward=True, scale: float = 1.0):
        """A plan built under the tuned options of this geometry (library defaults if it was never tuned)."""
        from . import _lib
        from .error import check
        from .plan import FftPlan

        lib = _lib.load()
        opts = self.options_for(shape, axes, kind, prec, forward)
        for k, v in opts.items():
            check(lib.sfc_planner_set_option(k.encode(), v.encode()))
        try:
            return FftPlan(shape, axes, kind, prec, forward, scale)
        finally:
            for k in opts:
                lib.sfc_planner_set_option(k.encode(), None)
