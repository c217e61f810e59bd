"""Developer benchmark: device-resident timing of the BASELINE configs with CUDA events."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scirs_b200 import FftPlan

HBM = 6553.9
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(plan, din, dout, iters=10, warm=3, do_flush=True):
    s = torch.cuda.current_stream()
    for _ in range(warm):
        plan.execute_device(din, dout, s.cuda_stream)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if do_flush: flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        plan.execute_device(din, dout, s.cuda_stream)
        e1.record(s)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]

def run(name, shape, axes, kind, prec, in_dt, out_dt, in_elems, out_elems, forward=True, iters=10):
    p = FftPlan(shape, axes, kind, prec, forward)
    din = torch.randn(in_elems * (2 if in_dt.is_complex else 1), device=dev, dtype=torch.float64 if prec == "f64" else torch.float32)
    dout = torch.empty(out_elems * (2 if out_dt.is_complex else 1), device=dev, dtype=din.dtype)
    med, best = timeit(p, din, dout, iters)
    alg = p.info["algorithmic_bytes"]; devb = p.info["device_bytes"]; fl = p.info["nominal_flops"]
    print(f"{name:34s} med {med:9.3f} ms best {best:9.3f} ms | alg {alg/1e9:7.3f} GB -> {alg/med/1e6:7.1f} GB/s ({alg/med/1e6/HBM:5.1%}) | dev {devb/1e9:7.3f} GB -> {devb/med/1e6:7.1f} GB/s | {fl/med/1e6:8.1f} GFLOP/s | launches {p.info['num_launches']}", flush=True)
    del din, dout
    return med

c128, f64, c64, f32 = torch.complex128, torch.float64, torch.complex64, torch.float32
which = sys.argv[1:] or ["all"]
def want(k): return "all" in which or k in which
if want("c2c4096"):
    run("c2c rows 65536x4096 f64", [65536, 4096], [1], "c2c", "f64", c128, c128, 65536 * 4096, 65536 * 4096)
    run("c2c rows 65536x4096 f32", [65536, 4096], [1], "c2c", "f32", c64, c64, 65536 * 4096, 65536 * 4096)
if want("rfft"):
    run("rfft 65536x4096 f64", [65536, 4096], [1], "r2c", "f64", f64, c128, 65536 * 4096, 65536 * 2049)
    run("irfft 65536x4096 f64", [65536, 4096], [1], "c2r", "f64", c128, f64, 65536 * 2049, 65536 * 4096)
    run("rfft 65536x4096 f32", [65536, 4096], [1], "r2c", "f32", f32, c64, 65536 * 4096, 65536 * 2049)
    run("irfft 65536x4096 f32", [65536, 4096], [1], "c2r", "f32", c64, f32, 65536 * 2049, 65536 * 4096)
if want("sizes"):
    for lg in range(4, 14):
        n = 1 << lg; b = (1 << 28) // n
        run(f"c2c rows {b}x{n} f64", [b, n], [1], "c2c", "f64", c128, c128, b * n, b * n, iters=5)
if want("fft2"):
    run("fft2 8192x8192 f64", [8192, 8192], [1, 0], "c2c", "f64", c128, c128, 8192 * 8192, 8192 * 8192)
    run("fft2 cols only 8192x8192", [8192, 8192], [0], "c2c", "f64", c128, c128, 8192 * 8192, 8192 * 8192)
if want("fft1m"):
    run("fft 2^20 batch 1 f64", [1, 1 << 20], [1], "c2c", "f64", c128, c128, 1 << 20, 1 << 20, iters=20)
    run("fft 2^20 batch 64 f64", [64, 1 << 20], [1], "c2c", "f64", c128, c128, 64 << 20, 64 << 20)
if want("fftn"):
    run("fftn 512^3 f64", [512, 512, 512], [0, 1, 2], "c2c", "f64", c128, c128, 512 ** 3, 512 ** 3)
    for a in (0, 1, 2):
        run(f"fftn 512^3 axis {a} only", [512, 512, 512], [a], "c2c", "f64", c128, c128, 512 ** 3, 512 ** 3, iters=5)
if want("fftn1024"):
    run("fftn 1024^3 f64", [1024, 1024, 1024], [0, 1, 2], "c2c", "f64", c128, c128, 1024 ** 3, 1024 ** 3, iters=3)
if want("blue"):
    run("bluestein 32x1000003 f64", [32, 1000003], [1], "c2c", "f64", c128, c128, 32 * 1000003, 32 * 1000003, iters=5)
    run("bluestein 32x1594323 f64", [32, 1594323], [1], "c2c", "f64", c128, c128, 32 * 1594323, 32 * 1594323, iters=5)
