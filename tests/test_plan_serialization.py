"""CPU: plan persistence mirror (plan_serialization.rs) — schema and bookkeeping, no device needed."""
import json
import os

import pytest


def test_database_round_trip_and_schema(tmp_path, build_artifacts):
    from scirs_b200.plan_serialization import PlanSerializationManager, PlanInfo

    path = os.path.join(tmp_path, "sub", "plans.json")
    m = PlanSerializationManager(path)
    assert not m.plan_exists(1024, True)                      # plan_serialization.rs:375-383 test_plan_serialization_basic
    info = m.create_plan_info(1024, True)
    assert info.arch_id.endswith("-sm_100a") and info.lib_version
    m.record_plan_usage(info, 5000)
    m.record_plan_usage(info, 7000)
    assert m.plan_exists(1024, True) and not m.plan_exists(1024, False)
    best = m.get_best_plan_metrics(1024, True)
    assert best[1].usage_count == 2 and best[1].avg_execution_ns == 6000
    m.save_database()
    d = json.load(open(path))
    # serde layout of the reference: list of [PlanInfo, PlanMetrics] pairs + stats + last_updated
    assert set(d) == {"plans", "stats", "last_updated"}
    assert set(d["plans"][0][0]) == {"size", "forward", "arch_id", "created_at", "lib_version"}
    assert set(d["plans"][0][1]) == {"avg_execution_ns", "usage_count", "last_used"}
    assert set(d["stats"]) == {"total_plans_created", "total_plans_loaded", "time_saved_ns"}
    m2 = PlanSerializationManager(path)                       # :386-409 persistence across managers
    assert m2.plan_exists(1024, True) and m2.get_best_plan_metrics(1024, True)[1].avg_execution_ns == 6000
    m2.set_enabled(False)
    assert not m2.plan_exists(1024, True) and m2.get_best_plan_metrics(1024, True) is None
    # a database written by the reference (CPU arch id) is readable and kept apart from GPU plans
    ref = {"plans": [[{"size": 64, "forward": True, "arch_id": "x86_64-avx2", "created_at": 1, "lib_version": "0.1.0-alpha.6"},
                      {"avg_execution_ns": 10, "usage_count": 3, "last_used": 2}]],
           "stats": {"total_plans_created": 1, "total_plans_loaded": 0, "time_saved_ns": 0}, "last_updated": 5}
    p3 = os.path.join(tmp_path, "ref.json")
    json.dump(ref, open(p3, "w"))
    m3 = PlanSerializationManager(p3)
    assert not m3.plan_exists(64, True) and m3.get_stats().total_plans_created == 1
    open(p3, "w").write("{not json")
    assert PlanSerializationManager(p3).get_stats().total_plans_created == 0   # :120-126 fallback to empty
