"""Auto-tuning mirror (auto_tuning.rs): bookkeeping on CPU, benchmarking + optimal dispatch on the GPU."""
import json
import os

import numpy as np
import pytest


def _mk(tmp_path):
    from scirs_b200.auto_tuning import AutoTuneConfig, AutoTuner, SizeRange, SizeStep

    cfg = AutoTuneConfig(sizes=SizeRange(16, 256, SizeStep.PowersOfTwo()), repetitions=3, warmup=1,
                         database_path=os.path.join(tmp_path, "db", "tune.json"))
    return AutoTuner(cfg)


def test_sizes_defaults_and_selection_rules(tmp_path, build_artifacts):
    from scirs_b200.auto_tuning import (AutoTuneConfig, BenchmarkResult, FftVariant, SizeRange, SizeStep, SystemInfo)

    d = AutoTuneConfig()  # auto_tuning.rs:73-89
    assert (d.sizes.min, d.sizes.max, d.repetitions, d.warmup) == (16, 8192, 10, 3)
    assert d.variants == [FftVariant.Standard, FftVariant.Cached] and d.database_path == ".fft_tuning_db.json"
    assert SizeRange(16, 128, SizeStep.PowersOfTwo()).sizes() == [16, 32, 64, 128]
    assert SizeRange(10, 50, SizeStep.Linear(20)).sizes() == [10, 30, 50]
    assert SizeRange(10, 100, SizeStep.Exponential(2.0)).sizes() == [10, 20, 40, 80]
    assert SizeRange(10, 100, SizeStep.Custom([5, 64, 99, 200])).sizes() == [64, 99]
    t = _mk(tmp_path)
    si = SystemInfo.detect()
    t.database.results += [BenchmarkResult(64, FftVariant.Standard, True, 900, 800, 1.0, si),
                           BenchmarkResult(64, FftVariant.Cached, True, 500, 400, 1.0, si),
                           BenchmarkResult(1024, FftVariant.Standard, True, 100, 90, 1.0, si),
                           BenchmarkResult(1024, FftVariant.Cached, True, 300, 90, 1.0, si)]
    t.update_best_algorithms()
    assert t.get_best_variant(64, True) == FftVariant.Cached          # exact
    assert t.get_best_variant(900, True) == FftVariant.Standard       # closest tuned size (1024)
    assert t.get_best_variant(100, True) == FftVariant.Cached         # closest tuned size (64)
    assert t.get_best_variant(64, False) == FftVariant.Standard       # nothing tuned for the inverse
    t.set_enabled(False)
    assert t.get_best_variant(64, True) == FftVariant.Standard and not t.is_enabled()
    t.set_enabled(True)
    t.save_database()
    raw = json.load(open(t.config.database_path))
    assert set(raw) == {"results", "last_updated", "best_algorithms"} and raw["results"][1]["variant"] == "Cached"
    t2 = _mk(tmp_path)                                                 # auto_tuning.rs tests: persistence
    assert t2.get_best_variant(64, True) == FftVariant.Cached and len(t2.database.results) == 4


@pytest.mark.gpu
def test_benchmarks_and_optimal_fft_on_gpu(tmp_path, build_artifacts):
    from scirs_b200.auto_tuning import FftVariant

    t = _mk(tmp_path)
    t.config.variants = [FftVariant.Standard, FftVariant.Cached, FftVariant.InPlace, FftVariant.SplitRadix]
    t.run_benchmarks()
    assert len(t.database.results) == 5 * 4 * 2 and os.path.exists(t.config.database_path)
    assert all(r.min_time_ns > 0 and r.avg_time_ns >= r.min_time_ns for r in t.database.results)
    x = np.random.default_rng(1).standard_normal(100) + 0j
    assert np.allclose(t.run_optimal_fft(x, 128, True), np.fft.fft(x, 128), atol=1e-10)
    assert np.allclose(t.run_optimal_fft(x, None, False), np.fft.ifft(x) * 100, atol=1e-10)  # unnormalised inverse


def test_gpu_plan_tuner_candidates_and_database(tmp_path):
    """The GPU planner tuner (round 2): candidate option sets per geometry and the JSON database — no GPU needed."""
    from scirs_b200.auto_tuning import GpuPlanTuner

    c = GpuPlanTuner.candidates([64, 1 << 20], [1])
    assert c[0] == {} and {"SFC_THREE_LEVEL_MIN": str(1 << 20)} in c and {"SFC_PIPE_LATE": "3"} in c
    assert {"SFC_BLUE_L1": "512"} in GpuPlanTuner.candidates([32, 1000003], [1])
    assert GpuPlanTuner.candidates([512, 512, 512], [0])[1:] == [{"SFC_PIPE_LATE": "3"}, {"SFC_COL_SMEM_KB": "40"}]
    db = str(tmp_path / "gpu_db.json")
    t = GpuPlanTuner(db)
    key = t.key([64, 1 << 20], [1], "c2c", "f64", True)
    t.entries[key] = {"options": {"SFC_PIPE_LATE": "3"}, "ms": 0.8, "default_ms": 0.9, "candidates": []}
    t.save()
    t2 = GpuPlanTuner(db)
    assert t2.options_for([64, 1 << 20], [1]) == {"SFC_PIPE_LATE": "3"} and t2.options_for([8, 8], [1]) == {}


def test_planner_options_round_trip():
    import ctypes as C

    from scirs_b200 import _lib

    lib = _lib.load()
    buf = C.create_string_buffer(64)
    assert lib.sfc_planner_get_option(b"SFC_TEST_OPTION_X", buf, 64) == 0
    assert lib.sfc_planner_set_option(b"SFC_TEST_OPTION_X", b"17") == 0
    assert lib.sfc_planner_get_option(b"SFC_TEST_OPTION_X", buf, 64) == 2 and buf.value == b"17"
    assert lib.sfc_planner_set_option(b"SFC_TEST_OPTION_X", None) == 0
    assert lib.sfc_planner_get_option(b"SFC_TEST_OPTION_X", buf, 64) == 0
    assert lib.sfc_planner_set_option(b"PATH", b"1") == _lib.SFC_ERR_VALUE


@pytest.mark.gpu
def test_gpu_plan_tuner_on_device(tmp_path):
    import numpy as np

    from scirs_b200.auto_tuning import GpuPlanTuner

    t = GpuPlanTuner(str(tmp_path / "db.json"), repetitions=5, warmup=2)
    e = t.tune([8, 1 << 17], [1])
    assert e["ms"] <= e["default_ms"] + 1e-9 and len(e["candidates"]) >= 4
    t.save()
    p = GpuPlanTuner(str(tmp_path / "db.json")).plan([8, 1 << 17], [1])
    x = np.random.default_rng(0).standard_normal((8, 1 << 17)) + 0j
    got = p.execute(x).reshape(8, -1)
    ref = np.fft.fft(x, axis=1)
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-12
