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

# consumer / multi-GPU plan features as well: fused DCT-II / III (rows, strided, sine), padded / cropped axes, scatter store
def run2(shape, axes, x, y, kind=_lib.SFC_C2C, flags=0, ail=0, aol=0, scale_dc=0.0, parts=0, outs=None):
    d = _lib.sfc_desc(); d.ndim = len(shape)
    for i, s in enumerate(shape): d.shape[i] = s
    d.naxes = len(axes)
    for i, a in enumerate(axes): d.axes[i] = a
    d.kind, d.prec, d.direction, d.flags, d.scale = kind, _lib.SFC_PREC_F64, 0, flags, 1.0
    d.axis_in_len, d.axis_out_len, d.scale_dc, d.scatter_parts = ail, aol, scale_dc, parts
    buf = C.create_string_buffer(8192)
    if parts:
        lib.emul_plan_run_scatter.argtypes = [C.POINTER(_lib.sfc_desc), C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_char_p, C.c_int]
        rc = lib.emul_plan_run_scatter(C.byref(d), x.ctypes.data_as(C.c_void_p), (C.c_void_p * parts)(*outs), parts, buf, len(buf))
    else:
        rc = lib.emul_plan_run(C.byref(d), x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), buf, len(buf))
    assert rc == 0, buf.value
    print(buf.value.decode().splitlines()[1][:100])

for shape, ax in (((64, 256), 1), ((512, 32), 0), ((16, 8192), 1)):
    for fl in (_lib.SFC_DESC_DCT2, _lib.SFC_DESC_DCT3, _lib.SFC_DESC_DCT2 | _lib.SFC_DESC_TRIG_SINE, _lib.SFC_DESC_DCT3 | _lib.SFC_DESC_TRIG_SINE):
        x = rng.standard_normal(shape); y = np.empty_like(x)
        run2(list(shape), [ax], x, y, _lib.SFC_R2C, fl, scale_dc=1.0)
x = rng.standard_normal((6, 100)) + 0j; y = np.empty((6, 60), dtype=complex)
run2([6, 256], [1], x, y, flags=_lib.SFC_DESC_AXIS_LEN, ail=100, aol=60)
x = rng.standard_normal((100, 12)) + 0j; y = np.empty((60, 12), dtype=complex)
run2([256, 12], [0], x, y, flags=_lib.SFC_DESC_AXIS_LEN, ail=100, aol=60)
x = rng.standard_normal((3, 70)) + 0j; y = np.empty((3, 90), dtype=complex)
run2([3, 90], [1], x, y, flags=_lib.SFC_DESC_AXIS_LEN, ail=70, aol=90)
P, s0, n1, n2 = 4, 4, 64, 32
recv = [np.empty((P * s0, n1 // P, n2), dtype=complex) for _ in range(P)]
x = rng.standard_normal((s0, n1, n2)) + 0j
run2([s0, n1, n2], [1], x, None, parts=P, outs=[recv[q].ctypes.data + 1 * s0 * (n1 // P) * n2 * 16 for q in range(P)])
print("asan feature run finished")
