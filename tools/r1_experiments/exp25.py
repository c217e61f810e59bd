import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scirs_b200 import FftPlan
dev = torch.device("cuda:0"); s = torch.cuda.current_stream()
for b, lg in ((4, 26), (2, 27), (1, 28), (1, 30)):
    n = 1 << lg
    x = torch.randn(b * n * 2, device=dev, dtype=torch.float64); y = torch.empty_like(x)
    p = FftPlan([b, n], [1])
    for _ in range(2): p.execute_device(x, y, s.cuda_stream)
    torch.cuda.synchronize(); ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); p.execute_device(x, y, s.cuda_stream); e1.record(s); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[1]; alg = p.info["algorithmic_bytes"]
    print(f"fft {b} x 2^{lg}: {t:8.3f} ms  alg {alg/1e9:.2f} GB -> {alg/t/1e6:6.0f} GB/s ({alg/t/1e6/6553.9:5.1%})  {5*b*n*lg/t/1e6:7.0f} GFLOP/s  passes {p.info['num_passes']}", flush=True)
    del x, y, p; torch.cuda.empty_cache()
