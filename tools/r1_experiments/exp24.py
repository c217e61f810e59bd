import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, scipy.fft as sf
from scirs_b200 import FftPlan
dev = torch.device("cuda:0"); s = torch.cuda.current_stream()
def run(shape, axis):
    tot = int(np.prod(shape))
    p = FftPlan(shape, [axis], "r2c", "f64", True, 1.0, dct2=True)
    x = torch.randn(tot, device=dev, dtype=torch.float64); y = torch.empty_like(x)
    for _ in range(3): p.execute_device(x, y, s.cuda_stream)
    torch.cuda.synchronize(); ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); p.execute_device(x, y, s.cuda_stream); e1.record(s); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[2]; byt = 2 * 8 * tot
    print(f"dct2 {shape} axis {axis}: {t:8.3f} ms {byt/t/1e6:7.0f} GB/s ({byt/t/1e6/6553.9:5.1%}) | {p.describe().splitlines()[1][10:100]}", flush=True)
run([16384, 16384], 1); run([1024, 1024 * 256], 0); run([256, 1 << 20], 0); run([4096, 65536], 0); run([64, 1024, 4096], 1)
