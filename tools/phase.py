"""Developer tool: per-phase clock breakdown of every launch of a plan (needs the -DSFC_PHASE_TIMING build)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SFC_PHASE_DBG"] = "1"
import torch
from scirs_b200 import FftPlan, _lib
lib = _lib.load()
dev = torch.device("cuda:0")
s = torch.cuda.current_stream()
def run(shape, axes, label, kind="c2c"):
    tot = 1
    for v in shape: tot *= v
    x = torch.randn(tot * 2, device=dev, dtype=torch.float64); y = torch.empty_like(x)
    p = FftPlan(shape, axes, kind, "f64", True)
    os.environ.pop("SFC_PHASE_DBG", None)
    for _ in range(2): p.execute_device(x, y, s.cuda_stream)
    torch.cuda.synchronize()
    os.environ["SFC_PHASE_DBG"] = "1"
    p.execute_device(x, y, s.cuda_stream)
    torch.cuda.synchronize()
    print("==", label); sys.stdout.flush()
    lib.sfc_debug_phase_dump()
    os.environ.pop("SFC_PHASE_DBG", None)
import ctypes as C
lib.sfc_debug_phase_dump.restype = None
run([16, 1000003], [1], "bluestein 16 x 1000003")
run([65536, 4096], [1], "c2c rows 65536 x 4096")
run([32768, 8192], [1], "c2c rows 32768 x 8192")
run([64, 1 << 20], [1], "fft 2^20 x 64")
run([512, 512, 512], [0, 1, 2], "fftn 512^3")
run([1048576, 256], [1], "c2c rows 1048576 x 256")
