"""BASELINE configs[0] as written: ONE 2^20-point c128 transform.  Device-resident latency per transform, back-to-back launches
and a CUDA graph of 10 transforms (knobs from the environment, e.g. SFC_COL_SMEM_KB=40 for narrower tiles)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scirs_b200 import FftPlan
n = 1 << 20
dev = torch.device("cuda:0")
plan = FftPlan([1, n], [1])
x = torch.randn(2 * n, dtype=torch.float64, device=dev); y = torch.empty_like(x)
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for _ in range(20): plan.execute_device(x, y, st.cuda_stream)
    st.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(500): plan.execute_device(x, y, st.cuda_stream)
    e1.record(st); st.synchronize()
    us = e0.elapsed_time(e1) / 500 * 1e3
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        for _ in range(10): plan.execute_device(x, y, st.cuda_stream)
    g.replay(); st.synchronize()
    e0.record(st)
    for _ in range(50): g.replay()
    e1.record(st); st.synchronize()
    gus = e0.elapsed_time(e1) / 500 * 1e3
knobs = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("SFC_"))
print(f"fft 2^20 batch 1: {us:.2f} us back-to-back, {gus:.2f} us in a CUDA graph (floor 10.24 us) | {knobs}")
print(plan.describe())
