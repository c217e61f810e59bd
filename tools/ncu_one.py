"""Run ONE plan a few times (for ncu captures).  usage: ncu_one.py <case>"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scirs_b200 import FftPlan
case = sys.argv[1]
dev = torch.device("cuda:0")
cases = {
    "c2c4096": ([65536, 4096], [1], "c2c", "f64", 65536 * 4096 * 2, 65536 * 4096 * 2),
    "c2c4096f32": ([65536, 4096], [1], "c2c", "f32", 65536 * 4096 * 2, 65536 * 4096 * 2),
    "rfft4096": ([65536, 4096], [1], "r2c", "f64", 65536 * 4096, 65536 * 2049 * 2),
    "irfft4096": ([65536, 4096], [1], "c2r", "f64", 65536 * 2049 * 2, 65536 * 4096),
    "c2c256": ([1048576, 256], [1], "c2c", "f64", (1 << 28) * 2, (1 << 28) * 2),
    "c2c8192": ([32768, 8192], [1], "c2c", "f64", (1 << 28) * 2, (1 << 28) * 2),
    "fftn512": ([512, 512, 512], [0, 1, 2], "c2c", "f64", 512 ** 3 * 2, 512 ** 3 * 2),
    "fft2_8192": ([8192, 8192], [1, 0], "c2c", "f64", 8192 ** 2 * 2, 8192 ** 2 * 2),
    "blue1m": ([16, 1000003], [1], "c2c", "f64", 16 * 1000003 * 2, 16 * 1000003 * 2),
    "fftn1024": ([1024, 1024, 1024], [0, 1, 2], "c2c", "f64", 1024 ** 3 * 2, 1024 ** 3 * 2),
    "fft1m64": ([64, 1 << 20], [1], "c2c", "f64", (64 << 20) * 2, (64 << 20) * 2),
}
cases["r3_13"] = ([32, 1594323], [1], "c2c", "f64", 32 * 1594323 * 2, 32 * 1594323 * 2)
cases["dct2"] = ([65536, 4096], [1], "r2c", "f64", 65536 * 4096, 65536 * 4096)
shape, axes, kind, prec, ni, no = cases[case]
dt = torch.float64 if prec == "f64" else torch.float32
p = FftPlan(shape, axes, kind, prec, True, dct2=(case == "dct2"))
din = torch.randn(ni, device=dev, dtype=dt)
dout = torch.empty(no, device=dev, dtype=dt)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    p.execute_device(din, dout, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print(p.describe())
