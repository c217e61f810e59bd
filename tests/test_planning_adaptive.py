"""planning_adaptive.rs: the reference's own unit tests (:318-360) restated, plus the switching rule (:202-238)."""
import numpy as np
import pytest

import scirs_b200 as sb


def test_adaptive_planner_basics():
    p = sb.AdaptivePlanner([16], True)  # planning_adaptive.rs:324-341
    assert p.current_strategy() == sb.PlanningStrategy.CacheFirst
    for _ in range(10):
        p.record_execution(100e-6)
    assert p.get_statistics()[sb.PlanningStrategy.CacheFirst][1] == 10
    assert abs(p.get_statistics()[sb.PlanningStrategy.CacheFirst][0] - 100e-6) < 1e-9


def test_strategy_switch_needs_threshold_samples_and_budget():
    cfg = sb.AdaptivePlanningConfig(evaluation_interval=0.0, min_samples=3, max_strategy_switches=1, improvement_threshold=1.1)
    p = sb.AdaptivePlanner([64], True, cfg)
    p._metrics[sb.PlanningStrategy.AutoTuned].record(95e-6)   # 5 % better: below the 10 % threshold
    for _ in range(4):
        p.record_execution(100e-6)
    assert p.current_strategy() == sb.PlanningStrategy.CacheFirst
    p._metrics[sb.PlanningStrategy.AlwaysNew].record(50e-6)   # 2x better
    p.record_execution(100e-6)
    assert p.current_strategy() == sb.PlanningStrategy.AlwaysNew
    p._metrics[sb.PlanningStrategy.SerializedFirst].record(1e-6)
    for _ in range(5):
        p.record_execution(50e-6)
    assert p.current_strategy() == sb.PlanningStrategy.AlwaysNew  # max_strategy_switches reached
    off = sb.AdaptivePlanner([64], True, sb.AdaptivePlanningConfig(enabled=False))
    off.record_execution(1.0)
    assert off.get_statistics()[sb.PlanningStrategy.CacheFirst][1] == 0


@pytest.mark.gpu
def test_adaptive_executor():
    ex = sb.AdaptiveExecutor([16], True)  # planning_adaptive.rs:343-360
    x = np.ones(16, dtype=np.complex128)
    out = np.zeros(16, dtype=np.complex128)
    for _ in range(5):
        ex.execute(x, out)
    assert ex.get_statistics()[ex.current_strategy()][1] >= 5
    assert abs(out[0] - 16.0) < 1e-12 and np.max(np.abs(out[1:])) < 1e-12


@pytest.mark.gpu
def test_fftn_memory_efficient_and_rfftn_optimized():
    from oracle import scirs2_fft_oracle as orc

    rng = np.random.default_rng(3)
    x = rng.standard_normal((6, 10, 16))
    # per-axis fft(&lane, None): pads every lane to the next power of two, keeps the first n bins (ndim_optimized.rs:60-86)
    ref = x.astype(np.complex128)
    for ax in (2, 0):
        n = ref.shape[ax]
        p = 1 << (n - 1).bit_length()
        ref = np.take(np.fft.fft(ref, p, axis=ax), range(n), axis=ax)
    got = sb.fftn_memory_efficient(x, [2, 0], 1.0)
    assert orc.rel_l2(got, ref) < 1e-12
    assert orc.rel_l2(sb.rfftn_optimized(x, None, [2, 0]), ref) < 1e-12
    with pytest.raises(sb.ValueError_):
        sb.fftn_memory_efficient(x, [3])
