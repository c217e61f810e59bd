import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scirs_b200 import FftPlan
dev = torch.device("cuda:0"); s = torch.cuda.current_stream()
def tables(n):
    i = np.arange(n); u = np.ones(n, dtype=np.complex128); w = np.exp(-1j * np.pi * i / (2 * n))
    return torch.from_numpy(u).to(dev), torch.from_numpy(w).to(dev)
def run(B, n):
    u, w = tables(n)
    p = FftPlan([B, 2 * n], [1], "c2c", "f64", True, 1.0, real_input=True, axis_in_len=n, axis_out_len=n, aux_in=u, aux_out=w, real_output=True)
    x = torch.randn(B * n, device=dev, dtype=torch.float64); y = torch.empty_like(x)
    for _ in range(3): p.execute_device(x, y, s.cuda_stream)
    torch.cuda.synchronize(); ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); p.execute_device(x, y, s.cuda_stream); e1.record(s); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[2]; byt = 2 * 8 * B * n
    import scipy.fft as sf
    ref = sf.dct(x[:n].cpu().numpy(), 2) / 2
    err = np.linalg.norm(y[:n].cpu().numpy() - ref) / np.linalg.norm(ref)
    print(f"dct2 {B}x{n}: {t:8.3f} ms  {byt/t/1e6:7.0f} GB/s ({byt/t/1e6/6553.9:5.1%})  err {err:.1e} | {p.describe().splitlines()[1][10:90]}", flush=True)
for n in (64, 256, 1024, 4096, 16384):
    run((1 << 28) // n, n)
