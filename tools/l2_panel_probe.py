"""Does an L2-resident working set make the two strided passes of a long column transform faster than HBM speed?
Column FFTs (axis 0, 8192 points: two four-step passes) over panels of 128 ... 8192 columns; small panels (in + work + out
<= ~100 MB) stay in the 126 MB L2 across repetitions, the big ones stream from HBM.  Prints effective GB/s (2 passes x r+w)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scirs_b200 import FftPlan
dev = torch.device("cuda:0"); s = torch.cuda.current_stream()
for cols in (128, 256, 512, 1024, 2048, 8192):
    n = 8192
    p = FftPlan([n, cols], [0])
    x = torch.randn(2 * n * cols, device=dev, dtype=torch.float64); y = torch.empty_like(x)
    for _ in range(5): p.execute_device(x, y, s.cuda_stream)
    reps = max(20, int(2000 * 256 / cols))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(reps): p.execute_device(x, y, s.cuda_stream)
    e1.record(s); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    byt = 2 * 2 * 16 * n * cols
    print(f"cols {cols:5d}: footprint {3 * 16 * n * cols / 1e6:7.1f} MB  {ms * 1e3:9.1f} us  {byt / ms / 1e6:8.1f} GB/s (two passes, read + write each) launches {p.info['num_launches']}", flush=True)
