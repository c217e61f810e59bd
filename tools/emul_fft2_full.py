"""One-off: the three-pass fft2 plan at FULL size (8192 x 8192) through the host emulation (tests/emul).  Minutes of CPU."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scirs_b200 import _lib
os.environ["SFC_FFT2_TILE2D"] = "1"
lib = C.CDLL(os.path.join(os.path.dirname(__file__), "..", "tests", "emul", "_build", "libplan_emul.so"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
rng = np.random.default_rng(3)
x = rng.standard_normal((n, 8192)) + 1j * rng.standard_normal((n, 8192))
y = np.zeros_like(x)
d = _lib.sfc_desc(); d.ndim = 2; d.shape[0] = n; d.shape[1] = 8192; d.naxes = 2; d.axes[0] = 0; d.axes[1] = 1
d.kind, d.prec, d.direction, d.flags, d.scale = _lib.SFC_C2C, _lib.SFC_PREC_F64, 0, 0, 1.0
buf = C.create_string_buffer(8192)
t0 = time.time()
rc = lib.emul_plan_run(C.byref(d), x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), buf, len(buf))
print("rc", rc, "seconds", round(time.time() - t0, 1)); print(buf.value.decode())
ref = np.fft.fft2(x)
print("rel-L2 vs numpy.fft.fft2:", np.linalg.norm(y - ref) / np.linalg.norm(ref))
