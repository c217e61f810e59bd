"""Plan persistence — mirror of scirs2-fft/src/plan_serialization.rs:49-352.

The JSON schema is the reference's (serde of `PlanDatabase`): ``plans`` is a list of
``[PlanInfo, PlanMetrics]`` pairs (plan_serialization.rs:19-45), ``stats`` a `PlanDatabaseStats`
and ``last_updated`` milliseconds since the epoch, so a database written by either side can be read
by the other.  The only extension is the architecture id: plans timed on this library carry the GPU
target after the host part (``x86_64-sm_100a``), which keeps them apart from CPU plans exactly the way
the reference keeps AVX and non-AVX plans apart (plan_serialization.rs:167-197).
"""
from __future__ import annotations

import json
import os
import platform
import threading
import time
from dataclasses import asdict, dataclass, field
from typing import Dict, Optional, Tuple

from . import _lib
from .error import IOError_, ValueError_

LIB_VERSION = "0.1.0-alpha.6"  # CARGO_PKG_VERSION of the reference this mirrors


def _now_ms() -> int:
    return int(time.time() * 1000)


@dataclass(frozen=True)
class PlanInfo:
    """plan_serialization.rs:49-60; identity = (size, forward, arch_id) (:63-70)."""
    size: int
    forward: bool
    arch_id: str
    created_at: int = 0
    lib_version: str = LIB_VERSION

    def key(self) -> Tuple[int, bool, str]:
        return (self.size, self.forward, self.arch_id)


@dataclass
class PlanMetrics:
    """plan_serialization.rs:86-93"""
    avg_execution_ns: int
    usage_count: int
    last_used: int


@dataclass
class PlanDatabaseStats:
    """plan_serialization.rs:97-104"""
    total_plans_created: int = 0
    total_plans_loaded: int = 0
    time_saved_ns: int = 0


@dataclass
class PlanDatabase:
    plans: Dict[Tuple[int, bool, str], Tuple[PlanInfo, PlanMetrics]] = field(default_factory=dict)
    stats: PlanDatabaseStats = field(default_factory=PlanDatabaseStats)
    last_updated: int = field(default_factory=_now_ms)

    def to_json(self) -> dict:
        return {"plans": [[asdict(i), asdict(m)] for i, m in self.plans.values()], "stats": asdict(self.stats),
                "last_updated": self.last_updated}

    @staticmethod
    def from_json(d: dict) -> "PlanDatabase":
        db = PlanDatabase()
        for info, met in d.get("plans", []):
            i = PlanInfo(int(info["size"]), bool(info["forward"]), str(info["arch_id"]), int(info.get("created_at", 0)),
                         str(info.get("lib_version", "")))
            db.plans[i.key()] = (i, PlanMetrics(int(met["avg_execution_ns"]), int(met["usage_count"]), int(met["last_used"])))
        db.stats = PlanDatabaseStats(**{k: int(v) for k, v in d.get("stats", {}).items()})
        db.last_updated = int(d.get("last_updated", _now_ms()))
        return db


class PlanSerializationManager:
    """plan_serialization.rs:107-333"""

    def __init__(self, db_path: str):
        self.db_path = str(db_path)
        self.enabled = True
        self._mu = threading.Lock()
        try:
            self.database = self._load_or_create(self.db_path)
        except Exception:  # :120-126: any failure falls back to an empty database
            self.database = PlanDatabase()

    @staticmethod
    def _load_or_create(path: str) -> PlanDatabase:
        if os.path.exists(path):
            try:
                with open(path, "r") as f:
                    raw = f.read()
            except OSError as e:
                raise IOError_(f"Failed to open plan database: {e}")
            try:
                return PlanDatabase.from_json(json.loads(raw))
            except Exception as e:
                raise ValueError_(f"Failed to parse plan database: {e}")
        parent = os.path.dirname(path)
        if parent:
            try:
                os.makedirs(parent, exist_ok=True)
            except OSError as e:
                raise IOError_(f"Failed to create directory for plan database: {e}")
        return PlanDatabase()

    @staticmethod
    def detect_arch_id() -> str:
        """Host architecture as in the reference (:167-197), then the GPU target this library compiles for."""
        m = platform.machine().lower()
        host = "x86_64" if m in ("x86_64", "amd64") else ("aarch64" if m in ("aarch64", "arm64") else f"unknown-{m}")
        return host + "-sm_100a"

    def create_plan_info(self, size: int, forward: bool) -> PlanInfo:
        return PlanInfo(int(size), bool(forward), self.detect_arch_id(), _now_ms(), LIB_VERSION)

    def plan_exists(self, size: int, forward: bool) -> bool:
        if not self.enabled:
            return False
        with self._mu:
            return (int(size), bool(forward), self.detect_arch_id()) in self.database.plans

    def record_plan_usage(self, plan_info: PlanInfo, execution_time_ns: int) -> None:
        """Running average exactly as :231-268 (integer truncation included)."""
        if not self.enabled:
            return
        save = False
        with self._mu:
            entry = self.database.plans.get(plan_info.key())
            if entry is None:
                entry = (plan_info, PlanMetrics(int(execution_time_ns), 0, _now_ms()))
                self.database.plans[plan_info.key()] = entry
            m = entry[1]
            m.usage_count += 1
            m.last_used = _now_ms()
            if m.usage_count > 1:
                m.avg_execution_ns = int((float(m.avg_execution_ns) * float(m.usage_count - 1) + float(execution_time_ns))
                                         / float(m.usage_count))
            else:
                m.avg_execution_ns = int(float(execution_time_ns))
            if self.database.last_updated + 60000 < _now_ms():
                save = True
                self.database.last_updated = _now_ms()
        if save:
            self.save_database()

    def save_database(self) -> None:
        if not self.enabled:
            return
        with self._mu:
            payload = json.dumps(self.database.to_json(), indent=2)
        try:
            with open(self.db_path, "w") as f:
                f.write(payload)
        except OSError as e:
            raise IOError_(f"Failed to create plan database file: {e}")

    def set_enabled(self, enabled: bool) -> None:
        self.enabled = bool(enabled)

    def get_best_plan_metrics(self, size: int, forward: bool) -> Optional[Tuple[PlanInfo, PlanMetrics]]:
        if not self.enabled:
            return None
        with self._mu:
            e = self.database.plans.get((int(size), bool(forward), self.detect_arch_id()))
            return None if e is None else (e[0], PlanMetrics(**asdict(e[1])))

    def get_stats(self) -> PlanDatabaseStats:
        with self._mu:
            return PlanDatabaseStats(**asdict(self.database.stats))


def create_and_time_plan(size: int, forward: bool):
    """plan_serialization.rs:335-352: build the plan, return it with the creation time in nanoseconds.
    The plan is a GPU plan (tables uploaded, kernels chosen); it needs a CUDA device."""
    from .plan import FftPlan

    t0 = time.perf_counter_ns()
    plan = FftPlan([int(size)], [0], "c2c", "f64", bool(forward))
    return plan, time.perf_counter_ns() - t0
