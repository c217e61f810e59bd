import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scirs_b200 import FftPlan
dev = torch.device("cuda:0"); s = torch.cuda.current_stream()
for prec, dt in (("f64", torch.float64), ("f32", torch.float32)):
    for n in (256, 512, 1024, 2048, 4096, 8192, 16384):
        b = (1 << 28) // n
        x = torch.randn(b * n, device=dev, dtype=dt); y = torch.empty(b * (n // 2 + 1) * 2, device=dev, dtype=dt)
        p = FftPlan([b, n], [1], "r2c", prec, True)
        for _ in range(3): p.execute_device(x, y, s.cuda_stream)
        torch.cuda.synchronize(); ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s); p.execute_device(x, y, s.cuda_stream); e1.record(s); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[2]; byt = (x.numel() + y.numel()) * x.element_size()
        print(f"rfft {prec} {b}x{n}: {t:7.3f} ms {byt/t/1e6:7.0f} GB/s ({byt/t/1e6/6553.9:5.1%}) | {p.describe().splitlines()[1][46:80]}", flush=True)
        del x, y
