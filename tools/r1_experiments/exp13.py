import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scirs_b200 import FftPlan
dev = torch.device("cuda:0")
s = torch.cuda.current_stream()
def run(shape, axes, label):
    tot = 1
    for v in shape: tot *= v
    x = torch.randn(tot * 2, device=dev, dtype=torch.float64); y = torch.empty_like(x)
    p = FftPlan(shape, axes, "c2c", "f64", True)
    for _ in range(3): p.execute_device(x, y, s.cuda_stream)
    torch.cuda.synchronize(); ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); p.execute_device(x, y, s.cuda_stream); e1.record(s); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[2]; byt = 2 * 16 * tot
    print(f"{label:34s} {t:8.3f} ms  compulsory {byt/t/1e6:7.0f} GB/s ({byt/t/1e6/6553.9:5.1%})  {p.describe().splitlines()[1][10:80]}", flush=True)
for n in (100, 1000, 1500, 3000, 4095, 5000, 10000, 100003):
    b = max(8, (1 << 27) // n)
    run([b, n], [1], f"bluestein rows {b}x{n}")
run([64, 1000, 1000], [1], "bluestein cols 64x1000x1000")
run([100, 100, 100], [0, 1, 2], "fftn 100^3")
run([1000, 1000], [1, 0], "fft2 1000x1000")
run([4096, 4096], [1, 0], "fft2 4096x4096")
run([256, 256, 256], [0, 1, 2], "fftn 256^3")
