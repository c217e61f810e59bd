"""Seeded sweep of plan shapes through the AddressSanitizer build of the host emulation: every plan the emulated planner can
build (f64; no copy / Hermitian-fill steps) is executed on exact-size arrays and compared with numpy.  Usage:
  LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python tools/emul_asan_sweep.py [count] [seed]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scirs_b200 import _lib
sub = os.environ.get("EMUL_BUILD", "asan" if os.environ.get("LD_PRELOAD") else ".")
lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "emul", "_build", sub, "libplan_emul.so"))
count = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)

def run(shape, axes, x, y, kind, inverse=0, scale=1.0):
    d = _lib.sfc_desc(); d.ndim = len(shape)
    for i, s in enumerate(shape): d.shape[i] = s
    d.naxes = len(axes)
    for i, a in enumerate(axes): d.axes[i] = a
    d.kind, d.prec, d.direction, d.flags, d.scale = kind, _lib.SFC_PREC_F64, inverse, 0, scale
    buf = C.create_string_buffer(16384)
    rc = lib.emul_plan_run(C.byref(d), x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), buf, len(buf))
    return rc, buf.value.decode()

sizes = [1, 2, 3, 4, 5, 7, 8, 12, 16, 17, 30, 32, 64, 100, 127, 128, 255, 256, 500, 512, 1000, 1024, 2048, 4096, 6000, 8192, 16384, 20000, 32768, 65536]
done = skipped = 0
worst = 0.0
while done < count:
    nd = int(rng.integers(1, 4))
    shape = [int(rng.choice(sizes[:14 if nd > 1 else len(sizes)])) for _ in range(nd)]
    if np.prod(shape) > 1 << 18:
        continue
    k = int(rng.integers(1, nd + 1))
    axes = [int(a) for a in rng.permutation(nd)[:k]]
    kind = int(rng.integers(0, 3))
    if kind == 0:
        inv = int(rng.integers(0, 2))
        x = rng.standard_normal(shape) + 1j * rng.standard_normal(shape); y = np.empty_like(x)
        rc, d = run(shape, axes, x, y, _lib.SFC_C2C, inv)
        ref = np.fft.ifftn(x, axes=axes) * np.prod([shape[a] for a in axes]) if inv else np.fft.fftn(x, axes=axes)
    elif kind == 1:
        x = rng.standard_normal(shape); hs = list(shape); hs[axes[-1]] = shape[axes[-1]] // 2 + 1
        y = np.empty(hs, dtype=np.complex128)
        rc, d = run(shape, axes, x, y, _lib.SFC_R2C)
        ref = np.fft.rfftn(x, axes=axes)
    else:
        hs = list(shape); hs[axes[-1]] = shape[axes[-1]] // 2 + 1
        full = rng.standard_normal(shape)
        x = np.ascontiguousarray(np.fft.rfftn(full, axes=axes)); y = np.empty(shape)
        rc, d = run(shape, axes, x, y, _lib.SFC_C2R)
        ref = full * np.prod([shape[a] for a in axes])
    if rc != 0:
        skipped += 1
        if "host emulation" not in d and "launch" not in d and "kernel" not in d.lower():
            print("plan refused:", shape, axes, kind, d[:100])
        continue
    e = np.linalg.norm((y - ref).ravel()) / max(np.linalg.norm(ref.ravel()), 1e-300)
    worst = max(worst, e)
    if e > 1e-12:
        print("MISMATCH", shape, axes, ["c2c", "r2c", "c2r"][kind], e, d.splitlines()[0])
    done += 1
print(f"{done} plans executed, {skipped} skipped (need un-emulated kernels), worst rel-L2 {worst:.2e}")
