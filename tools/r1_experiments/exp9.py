import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scirs_b200 import FftPlan
dev = torch.device("cuda:0")
n = 1024
x = torch.randn(n * n * n * 2, device=dev, dtype=torch.float64)
y = torch.empty_like(x)
s = torch.cuda.current_stream()
for ax in (0, 1, 2):
    p = FftPlan([n, n, n], [ax], "c2c", "f64", True)
    for _ in range(2): p.execute_device(x, y, s.cuda_stream)
    torch.cuda.synchronize()
    ts = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); p.execute_device(x, y, s.cuda_stream); e1.record(s); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    print(f"1024^3 axis {ax}: {t:.3f} ms  {2 * 16 * n**3 / t / 1e6:.0f} GB/s ({2 * 16 * n**3 / t / 1e6 / 6553.9:.1%})  {p.describe().splitlines()[1][:90]}", flush=True)
