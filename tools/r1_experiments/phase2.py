"""Developer tool: per-phase clocks of the real-transform / DCT kernels (needs the -DSFC_PHASE_TIMING build)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scirs_b200 import FftPlan, _lib
lib = _lib.load(); lib.sfc_debug_phase_dump.restype = None
dev = torch.device("cuda:0"); s = torch.cuda.current_stream()
def run(label, shape, kind, nin, nout, **kw):
    x = torch.randn(nin, device=dev, dtype=torch.float64); y = torch.empty(nout, device=dev, dtype=torch.float64)
    p = FftPlan(shape, [1], kind, "f64", kind != "c2r", **kw)
    for _ in range(2): p.execute_device(x, y, s.cuda_stream)
    torch.cuda.synchronize()
    os.environ["SFC_PHASE_DBG"] = "1"
    p.execute_device(x, y, s.cuda_stream); torch.cuda.synchronize()
    print("==", label); sys.stdout.flush(); lib.sfc_debug_phase_dump(); os.environ.pop("SFC_PHASE_DBG", None)
B, n = 65536, 4096
run("rfft", [B, n], "r2c", B * n, B * (n // 2 + 1) * 2)
run("irfft", [B, n], "c2r", B * (n // 2 + 1) * 2, B * n)
run("dct2", [B, n], "r2c", B * n, B * n, dct2=True)
run("c2c", [B, n], "c2c", B * n * 2, B * n * 2)
run("rfft 1024", [B * 4, 1024], "r2c", B * n, B * 4 * 513 * 2)
