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
