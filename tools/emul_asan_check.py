"""Run the experimental plans through the AddressSanitizer build of the host emulation (out-of-bounds accesses of the
kernels or the planner's work areas show up as ASan reports).  Usage:
  LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python tools/emul_asan_check.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scirs_b200 import _lib
os.environ["SFC_FFT2_TILE2D"] = "1"
lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "emul", "_build", "asan", "libplan_emul.so"))
rng = np.random.default_rng(0)

def run(shape, axes, x, y, kind=_lib.SFC_C2C, flags=0, inverse=0):
    d = _lib.sfc_desc(); d.ndim = len(shape)
    for i, s in enumerate(shape): d.shape[i] = s
    d.naxes = len(axes)
    for i, a in enumerate(axes): d.axes[i] = a
    d.kind, d.prec, d.direction, d.flags, d.scale = kind, _lib.SFC_PREC_F64, inverse, flags, 1.0
    buf = C.create_string_buffer(8192)
    rc = lib.emul_plan_run(C.byref(d), x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), buf, len(buf))
    assert rc == 0, buf.value
    return buf.value.decode().splitlines()[1]

# exact-size heap arrays so that any access past the end lands in a red zone
for shape, ax in (((32, 128), 1), ((4, 2048), 1), ((256, 64), 0), ((4, 512, 16), 1)):
    for fl in (_lib.SFC_DESC_DCT4, _lib.SFC_DESC_DCT4 | _lib.SFC_DESC_TRIG_SINE):
        x = rng.standard_normal(shape); y = np.empty_like(x)
        print(run(list(shape), [ax], x, y, _lib.SFC_R2C, fl))
for inv in (0, 1):
    x = rng.standard_normal((256, 8192)) + 1j * rng.standard_normal((256, 8192)); y = np.empty_like(x)
    print(run([256, 8192], [0, 1], x, y, inverse=inv))
    ref = np.fft.ifft2(x) * x.size if inv else np.fft.fft2(x)
    print("  rel-L2", np.linalg.norm(y - ref) / np.linalg.norm(ref))
print("asan run finished")
