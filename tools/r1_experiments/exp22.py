import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, scipy.fft as sf
from scirs_b200 import FftPlan
dev = torch.device("cuda:0"); s = torch.cuda.current_stream()
def run(B, n, ortho=False):
    p = FftPlan([B, n], [1], "r2c", "f64", True, (2.0 / n) ** 0.5 if ortho else 1.0, dct2=True, dct2_ortho=ortho)
    x = torch.randn(B * n, device=dev, dtype=torch.float64); y = torch.empty_like(x)
    for _ in range(3): p.execute_device(x, y, s.cuda_stream)
    torch.cuda.synchronize(); ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); p.execute_device(x, y, s.cuda_stream); e1.record(s); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[2]; byt = 2 * 8 * B * n
    xs = x[:4 * n].cpu().numpy().reshape(4, n)
    ref = sf.dct(xs, 2, norm="ortho") if ortho else sf.dct(xs, 2) / 2
    got = y[:4 * n].cpu().numpy().reshape(4, n)
    err = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    print(f"dct2 {B}x{n} ortho={ortho}: {t:8.3f} ms  {byt/t/1e6:7.0f} GB/s ({byt/t/1e6/6553.9:5.1%})  err {err:.1e} | {p.describe().splitlines()[1][10:90]}", flush=True)
for n in (128, 256, 1024, 4096, 8192, 16384):
    run((1 << 28) // n, n, n == 1024)
