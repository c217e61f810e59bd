"""Copy-engine peer copies under load (one process, 2+ GPUs): how fast does cudaMemcpyPeerAsync move a slab block to the
peer(s), alone and while an HBM-bound transform pass runs on both GPUs?  Decides whether a copy-engine exchange (compute
passes store locally, DMA does the all-to-all) can beat the fused store-over-NVLink scatter pass."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scirs_b200 import FftPlan
import scirs_b200 as sb
from scirs_b200 import _lib

ng = torch.cuda.device_count()
lib = _lib.load()
MB = 1 << 20
nbytes = 256 * MB  # one direction, per peer pair
src = [torch.empty(nbytes // 8, dtype=torch.float64, device=f"cuda:{g}").normal_() for g in range(ng)]
dst = [[torch.empty(nbytes // 8 // max(ng - 1, 1), dtype=torch.float64, device=f"cuda:{q}") for q in range(ng)] for g in range(ng)]
cs = [[torch.cuda.Stream(device=f"cuda:{g}") for q in range(ng)] for g in range(ng)]
# an HBM-bound pass per GPU: 512^3 / ng slab, rows
plans, bufs, ks = [], [], []
for g in range(ng):
    torch.cuda.set_device(g)
    sb.error.check(lib.sfc_init(g))
    x = torch.randn(2 * 65536 * 4096 // 4, dtype=torch.float64, device=f"cuda:{g}")
    y = torch.empty_like(x)
    p = FftPlan([65536 // 4, 4096], [1])
    plans.append(p); bufs.append((x, y)); ks.append(torch.cuda.Stream(device=f"cuda:{g}"))

def copies():
    for g in range(ng):
        torch.cuda.set_device(g)
        per = src[g].numel() // max(ng - 1, 1)
        k = 0
        for q in range(ng):
            if q == g: continue
            with torch.cuda.stream(cs[g][q]):
                dst[g][q].copy_(src[g][k * per:(k + 1) * per], non_blocking=True)
            k += 1

def kernels(n):
    for g in range(ng):
        torch.cuda.set_device(g)
        for _ in range(n):
            plans[g].execute_device(bufs[g][0], bufs[g][1], ks[g].cuda_stream)

def sync():
    for g in range(ng):
        torch.cuda.synchronize(g)

for load in (0, 1):
    for _ in range(2):
        copies(); kernels(2 * load); sync()
    t0 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        if load: kernels(2)
        copies()
    sync()
    dt = (time.perf_counter() - t0) / reps
    print(f"{ng} GPUs, every GPU sends {nbytes / MB:.0f} MiB split over its peers ({'with' if load else 'without'} 2 x 2.1 GB transform passes per GPU alongside): "
          f"{dt * 1e3:.3f} ms per round -> {nbytes / dt / 1e9:.1f} GB/s per GPU per direction", flush=True)
if True:
    kernels(2); sync()
    t0 = time.perf_counter()
    for _ in range(10): kernels(2)
    sync()
    print(f"the 2 transform passes alone: {(time.perf_counter() - t0) / 10 * 1e3:.3f} ms per round")
