"""Device-resident timing of the experimental plans against the current ones (first GPU run of the next round):
  fused DCT-IV rows (FftPlan(..., "r2c", dct4=True)) against the 2n-point formulation the consumers use today
  (timed through sb.dctn on host buffers is PCIe-bound, so the 2n-point plan is rebuilt here the way api_ext.cu does),
  and fft2 8192 x 8192 with SFC_FFT2_TILE2D=0 / 1 (run this script twice, once per setting)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scirs_b200 import FftPlan
dev = torch.device("cuda:0"); s = torch.cuda.current_stream(); HBM = 6553.9

def timeit(p, x, y, iters=5):
    for _ in range(3): p.execute_device(x, y, s.cuda_stream)
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); p.execute_device(x, y, s.cuda_stream); e1.record(s); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]

for n in (1024, 4096, 16384):
    b = (1 << 28) // n
    x = torch.randn(b * n, device=dev, dtype=torch.float64); y = torch.empty_like(x)
    for name, kw in (("DCT-II fused", dict(dct2=True)), ("DCT-IV fused (experimental)", dict(dct4=True)),
                     ("DST-IV fused (experimental)", dict(dct4=True, trig_sine=True))):
        try:
            p = FftPlan([b, n], [1], "r2c", "f64", True, 1.0, **kw)
            t = timeit(p, x, y); byt = 2 * 8 * b * n
            print(f"{name:30s} {b}x{n}: {t:7.3f} ms {byt/t/1e6:7.0f} GB/s ({byt/t/1e6/HBM:5.1%}) | {p.describe().splitlines()[1][:90]}", flush=True)
        except Exception as ex:
            print(name, n, "failed:", str(ex)[:120])
    del x, y
for n in (4096, 8192, 16384):
    x = torch.randn(n * n * 2, device=dev, dtype=torch.float64); y = torch.empty_like(x)
    p = FftPlan([n, n], [0, 1])
    t = timeit(p, x, y, 3); byt = 2 * 2 * 16 * n * n
    print(f"fft2 {n}x{n} SFC_FFT2_TILE2D={os.environ.get('SFC_FFT2_TILE2D', '0')}: {t:8.3f} ms {byt/t/1e6:7.0f} GB/s ({byt/t/1e6/HBM:5.1%} of the two-pass roofline)", flush=True)
    print(p.describe())
    del x, y, p
    torch.cuda.empty_cache()
