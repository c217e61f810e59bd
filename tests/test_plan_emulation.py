"""CPU: the PLANNER (scirs_b200/csrc/plan.cu) and the f64 kernel translation units compiled for the host
(-DSFC_HOST_EMUL, tests/emul/cuda_runtime.h) and driven through Plan::create / Plan::exec on host arrays.

Purpose: check planner plumbing + kernel index logic of code that has not run on a GPU yet (the fused DCT-IV flavour and
the three-pass fft2 plan), after showing on GPU-validated plans that the emulation reproduces them.  Test infrastructure
only — see tests/emul/cuda_runtime.h; the product library is untouched and still has no CPU path.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL = os.path.join(ROOT, "tests", "emul")
CSRC = os.path.join(ROOT, "scirs_b200", "csrc")
UNITS = ["plan", "aux_kernels", "kernels_f64_small", "kernels_f64_mid", "kernels_f64_big", "kernels_f64_real", "kernels_f64_dbl_a",
         "kernels_f64_dbl_b", "kernels_dct", "kernels_r3"]


@pytest.fixture(scope="module")
def emul():
    from scirs_b200 import _lib

    bdir = os.path.join(EMUL, "_build", "obj")
    os.makedirs(bdir, exist_ok=True)
    out = os.path.join(EMUL, "_build", "libplan_emul.so")
    hdrs = [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".h", ".cuh"))] + [os.path.join(EMUL, "cuda_runtime.h")]
    newest_hdr = max(os.path.getmtime(h) for h in hdrs)
    flags = ["-std=c++17", "-O1", "-fPIC", "-pthread", "-DSFC_HOST_EMUL", "-I" + EMUL, "-I" + CSRC, "-I" + os.path.join(ROOT, "include")]
    procs, objs = [], []
    for u in UNITS + ["plan_emul"]:
        src = os.path.join(EMUL, u + ".cpp") if u == "plan_emul" else os.path.join(CSRC, u + ".cu")
        obj = os.path.join(bdir, u + ".o")
        objs.append(obj)
        if not os.path.exists(obj) or os.path.getmtime(obj) < max(newest_hdr, os.path.getmtime(src)):
            procs.append(subprocess.Popen(["g++", "-x", "c++"] + flags + ["-c", src, "-o", obj]))
    for pr in procs:
        assert pr.wait() == 0, "host compilation of the planner / kernels failed"
    if procs or not os.path.exists(out):
        subprocess.run(["g++", "-shared", "-pthread", "-o", out] + objs, check=True)
    old_knob = os.environ.get("SFC_FFT2_TILE2D")
    os.environ["SFC_FFT2_TILE2D"] = "1"  # read once by the emulated planner, at its first 2-D plan
    lib = C.CDLL(out)
    lib.emul_plan_run.argtypes = [C.POINTER(_lib.sfc_desc), C.c_void_p, C.c_void_p, C.c_char_p, C.c_int]

    def run(shape, axes, x, out_arr, kind=_lib.SFC_C2C, inverse=False, scale=1.0, flags=0, axis_in_len=0, axis_out_len=0,
            scale_dc=0.0):
        d = _lib.sfc_desc()
        d.axis_in_len, d.axis_out_len, d.scale_dc = axis_in_len, axis_out_len, scale_dc
        d.ndim = len(shape)
        for i, s in enumerate(shape):
            d.shape[i] = s
        d.naxes = len(axes)
        for i, a in enumerate(axes):
            d.axes[i] = a
        d.kind, d.prec, d.direction, d.flags, d.scale = kind, _lib.SFC_PREC_F64, int(inverse), flags, scale
        buf = C.create_string_buffer(8192)
        rc = lib.emul_plan_run(C.byref(d), x.ctypes.data_as(C.c_void_p), out_arr.ctypes.data_as(C.c_void_p), buf, len(buf))
        return rc, buf.value.decode()

    def run_scatter(shape, axes, x, outs, parts, pitch=0, inverse=False, scale=1.0):
        d = _lib.sfc_desc()
        d.ndim = len(shape)
        for i, s in enumerate(shape):
            d.shape[i] = s
        d.naxes = len(axes)
        for i, a in enumerate(axes):
            d.axes[i] = a
        d.kind, d.prec, d.direction, d.flags, d.scale, d.scatter_parts = _lib.SFC_C2C, _lib.SFC_PREC_F64, int(inverse), 0, scale, parts
        d.scatter_pitch = pitch
        ptrs = (C.c_void_p * parts)(*outs)
        buf = C.create_string_buffer(8192)
        rc = lib.emul_plan_run_scatter(C.byref(d), x.ctypes.data_as(C.c_void_p), ptrs, parts, buf, len(buf))
        return rc, buf.value.decode()

    def run_windows(shape, axes, x, out_arr, outs, parts, row_lanes, nwin, pitch=0, inverse=False, scale=1.0):
        d = _lib.sfc_desc()
        d.ndim = len(shape)
        for i, s in enumerate(shape):
            d.shape[i] = s
        d.naxes = len(axes)
        for i, a in enumerate(axes):
            d.axes[i] = a
        d.kind, d.prec, d.direction, d.flags, d.scale, d.scatter_parts = _lib.SFC_C2C, _lib.SFC_PREC_F64, int(inverse), 0, scale, parts
        d.scatter_pitch = pitch
        ptrs = (C.c_void_p * max(parts, 1))(*outs) if parts else None
        buf = C.create_string_buffer(8192)
        rc = lib.emul_plan_run_windows(C.byref(d), x.ctypes.data_as(C.c_void_p),
                                       out_arr.ctypes.data_as(C.c_void_p) if out_arr is not None else None, ptrs, parts,
                                       row_lanes, nwin, buf, len(buf))
        return rc, buf.value.decode()

    run.scatter = run_scatter
    run.windows = run_windows
    lib.emul_plan_run_windows.argtypes = [C.POINTER(_lib.sfc_desc), C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_longlong,
                                          C.c_int, C.c_char_p, C.c_int]
    lib.emul_plan_run_scatter.argtypes = [C.POINTER(_lib.sfc_desc), C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_char_p, C.c_int]
    yield run
    if old_knob is None:
        os.environ.pop("SFC_FFT2_TILE2D", None)
    else:
        os.environ["SFC_FFT2_TILE2D"] = old_knob


def rel(a, b):
    return np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel())


def test_emulated_planner_reproduces_validated_plans(emul):
    """Plans whose GPU results are already pinned by tests/test_gpu_parity.py: single-pass rows and columns, a 2-D transform,
    a four-step row, Bluestein, fused rfft / irfft, the fused DCT-II."""
    from scirs_b200 import _lib
    from oracle import consumers_oracle as co

    rng = np.random.default_rng(1)
    c = lambda *s: rng.standard_normal(s) + 1j * rng.standard_normal(s)
    x = c(8, 1024); y = np.zeros_like(x)
    rc, d = emul([8, 1024], [1], x, y); assert rc == 0, d
    assert rel(y, np.fft.fft(x, axis=1)) < 1e-14
    x = c(256, 24); y = np.zeros_like(x)
    rc, d = emul([256, 24], [0], x, y, inverse=True, scale=1 / 256); assert rc == 0, d
    assert rel(y, np.fft.ifft(x, axis=0)) < 1e-14
    x = c(64, 128); y = np.zeros_like(x)
    rc, d = emul([64, 128], [0, 1], x, y); assert rc == 0, d
    assert rel(y, np.fft.fft2(x)) < 1e-14
    x = c(2, 1 << 15); y = np.zeros_like(x)
    rc, d = emul([2, 1 << 15], [1], x, y); assert rc == 0 and "four-step" in d, d
    assert rel(y, np.fft.fft(x, axis=1)) < 1e-14
    x = c(3, 1000); y = np.zeros_like(x)
    rc, d = emul([3, 1000], [1], x, y); assert rc == 0, d
    assert rel(y, np.fft.fft(x, axis=1)) < 1e-13
    xr = rng.standard_normal((4, 2048)); yh = np.zeros((4, 1025), dtype=np.complex128)
    rc, d = emul([4, 2048], [1], xr, yh, kind=_lib.SFC_R2C); assert rc == 0, d
    assert rel(yh, np.fft.rfft(xr, axis=1)) < 1e-14
    back = np.zeros_like(xr)
    rc, d = emul([4, 2048], [1], yh, back, kind=_lib.SFC_C2R, scale=1 / 2048); assert rc == 0, d
    assert rel(back, xr) < 1e-14
    xd = rng.standard_normal((64, 128)); yd = np.zeros_like(xd)
    rc, d = emul([64, 128], [1], xd, yd, kind=_lib.SFC_R2C, flags=_lib.SFC_DESC_DCT2); assert rc == 0 and "fused DCT-II" in d, d
    assert rel(yd, co.dctn(xd, 2, None, [1])) < 1e-13


@pytest.mark.parametrize("shape,axis", [((32, 128), 1), ((8, 1024), 1), ((2, 16384), 1), ((256, 64), 0), ((4, 512, 16), 1)])
def test_fused_dct4_plan(emul, shape, axis):
    """SFC_DESC_DCT4 through Plan::create -> add_dct4 -> TM_FAST_DCT4, cosine and sine, against the literal sums."""
    from scirs_b200 import _lib
    from oracle import consumers_oracle as co

    rng = np.random.default_rng(sum(shape))
    x = rng.standard_normal(shape)
    for sine in (False, True):
        y = np.zeros_like(x)
        rc, d = emul(list(shape), [axis], x, y, kind=_lib.SFC_R2C, scale=0.5,
                     flags=_lib.SFC_DESC_DCT4 | (_lib.SFC_DESC_TRIG_SINE if sine else 0))
        assert rc == 0 and ("fused DST-IV" if sine else "fused DCT-IV") in d, d
        if shape[axis] <= 4096:
            ref = 0.5 * (co.dstn(x, 4, None, [axis]) / 2.0 if sine else co.dctn(x, 4, None, [axis]))
        else:  # the literal O(n^2) sums lose 1e-12 themselves at this length: scipy's type IV is twice the plain sum
            import scipy.fft as sf

            ref = 0.25 * (sf.dst(x, 4, axis=axis) if sine else sf.dct(x, 4, axis=axis))
        assert rel(y, ref) < 1e-12


@pytest.mark.parametrize("R,Cn", [(256, 8192), (512, 4096), (256, 16384)])
def test_fft2_three_pass_plan(emul, R, Cn):
    """SFC_FFT2_TILE2D=1: Plan::create -> add_fft2_three_pass (the plan takes any power-of-two row count from 256 with 4096,
    8192 or 16384 columns; 8192 x 8192 runs the same three kernels with a longer first pass)."""
    rng = np.random.default_rng(2)
    x = rng.standard_normal((R, Cn)) + 1j * rng.standard_normal((R, Cn))
    y = np.zeros_like(x)
    rc, d = emul([R, Cn], [0, 1], x, y, scale=0.125)
    assert rc == 0 and d.count("2-D three-pass") == 3, d
    assert rel(y, 0.125 * np.fft.fft2(x)) < 1e-13
    back = np.zeros_like(x)
    rc, d = emul([R, Cn], [1, 0], y, back, inverse=True, scale=8.0 / (R * Cn))
    assert rc == 0 and d.count("2-D three-pass") == 3, d
    assert rel(back, x) < 1e-13


def test_seeded_plan_sweep(emul):
    """40 random plans (c2c forward / inverse, r2c, c2r; 1 to 3 dimensions; power-of-two, Bluestein and tiny extents) through
    the emulated planner + kernels against numpy (tools/emul_asan_sweep.py, the same sweep that is run under
    AddressSanitizer by hand)."""
    import sys

    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "emul_asan_sweep.py"), "40", "7"], capture_output=True,
                       text=True, timeout=900, env={k: v for k, v in os.environ.items() if k != "LD_PRELOAD"})
    assert r.returncode == 0, r.stdout + r.stderr
    assert "40 plans executed" in r.stdout and "MISMATCH" not in r.stdout and "plan refused" not in r.stdout, r.stdout


KNOB_SCRIPT = r"""
import ctypes as C, os, sys
sys.path.insert(0, %r)
import numpy as np
from scirs_b200 import _lib
lib = C.CDLL(%r)
rng = np.random.default_rng(4)
def run(shape, axes, x, y, kind=_lib.SFC_C2C, inverse=0, scale=1.0):
    d = _lib.sfc_desc(); d.ndim = len(shape)
    for i, s in enumerate(shape): d.shape[i] = s
    d.naxes = len(axes)
    for i, a in enumerate(axes): d.axes[i] = a
    d.kind, d.prec, d.direction, d.flags, d.scale = kind, _lib.SFC_PREC_F64, inverse, 0, scale
    buf = C.create_string_buffer(16384)
    rc = lib.emul_plan_run(C.byref(d), x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), buf, len(buf))
    assert rc == 0, buf.value
    return buf.value.decode()
for shape in ([2, 1 << 16], [1, 1 << 18]):
    x = rng.standard_normal(shape) + 1j * rng.standard_normal(shape); y = np.empty_like(x)
    d = run(shape, [1], x, y)
    assert WANT in d, d
    e = np.linalg.norm(y - np.fft.fft(x, axis=1)) / np.linalg.norm(y)
    back = np.empty_like(x)
    run(shape, [1], y, back, inverse=1, scale=1.0 / shape[1])
    e2 = np.linalg.norm(back - x) / np.linalg.norm(x)
    assert e < 1e-14 and e2 < 1e-14, (shape, e, e2)
n = 100003
x = rng.standard_normal((2, n)) + 1j * rng.standard_normal((2, n)); y = np.empty_like(x)
d = run([2, n], [1], x, y)
assert "Bluestein" in d or "chirp" in d, d
assert np.linalg.norm(y - np.fft.fft(x, axis=1)) / np.linalg.norm(y) < 1e-13
print("knob plans ok")
"""


@pytest.mark.parametrize("env,want", [({"SFC_THREE_LEVEL_MIN": "16384"}, "three-level"), ({"SFC_THREE_LEVEL_MIN": "0"}, "four-step")])
def test_long_rows_and_bluestein_through_the_emulation(emul, env, want):
    """Long rows in both decompositions (the knobs are read once per process, hence the subprocess) and a three-pass
    Bluestein length, forward and inverse, against numpy."""
    import sys

    script = (KNOB_SCRIPT % (ROOT, os.path.join(EMUL, "_build", "libplan_emul.so"))).replace("WANT", repr(want))
    r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=900, env=dict(os.environ, **env))
    assert r.returncode == 0 and "knob plans ok" in r.stdout, r.stdout + r.stderr


def test_consumer_plan_features(emul):
    """Descriptor features the consumers rely on (api_ext.cu): fused DCT-II / DCT-III and their sine twins on rows and on a
    strided axis, and a padded / cropped axis (SFC_DESC_AXIS_LEN: zero-padding on load, crop on store)."""
    from scirs_b200 import _lib
    from oracle import consumers_oracle as co

    rng = np.random.default_rng(6)
    for shape, axis in (((64, 256), 1), ((512, 32), 0)):
        x = rng.standard_normal(shape)
        n = shape[axis]
        for sine in (False, True):
            fl = _lib.SFC_DESC_TRIG_SINE if sine else 0
            y = np.zeros_like(x)
            rc, d = emul(list(shape), [axis], x, y, kind=_lib.SFC_R2C, flags=_lib.SFC_DESC_DCT2 | fl)
            assert rc == 0, d
            ref = co.dstn(x, 2, None, [axis]) if sine else co.dctn(x, 2, None, [axis])
            assert rel(y, ref) < 1e-13, (shape, axis, sine)
            # type III with the scales of api_ext.cu trig_axis: dct3 forward (None) = kernel(scale 1/n, dc 1); dst3 forward = (0.25, dc 2)
            y = np.zeros_like(x)
            rc, d = emul(list(shape), [axis], x, y, kind=_lib.SFC_R2C, flags=_lib.SFC_DESC_DCT3 | fl,
                         scale=0.25 if sine else 1.0 / n, scale_dc=2.0 if sine else 1.0)
            assert rc == 0, d
            ref = co.dstn(x, 3, None, [axis]) if sine else co.dctn(x, 3, None, [axis])
            assert rel(y, ref) < 1e-13, (shape, axis, sine, "III")
    # 100 input samples zero-padded to a 256-point transform, first 60 bins kept
    x = rng.standard_normal((6, 100)) + 1j * rng.standard_normal((6, 100))
    y = np.zeros((6, 60), dtype=np.complex128)
    rc, d = emul([6, 256], [1], x, y, flags=_lib.SFC_DESC_AXIS_LEN, axis_in_len=100, axis_out_len=60)
    assert rc == 0, d
    assert rel(y, np.fft.fft(x, 256, axis=1)[:, :60]) < 1e-14
    # the same on a strided axis, and with a Bluestein length
    x = rng.standard_normal((100, 12)) + 1j * rng.standard_normal((100, 12))
    y = np.zeros((60, 12), dtype=np.complex128)
    rc, d = emul([256, 12], [0], x, y, flags=_lib.SFC_DESC_AXIS_LEN, axis_in_len=100, axis_out_len=60)
    assert rc == 0, d
    assert rel(y, np.fft.fft(x, 256, axis=0)[:60]) < 1e-14
    x = rng.standard_normal((3, 70)) + 1j * rng.standard_normal((3, 70))
    y = np.zeros((3, 90), dtype=np.complex128)
    rc, d = emul([3, 90], [1], x, y, flags=_lib.SFC_DESC_AXIS_LEN, axis_in_len=70, axis_out_len=90)
    assert rc == 0, d
    assert rel(y, np.fft.fft(x, 90, axis=1)) < 1e-13


@pytest.mark.parametrize("P", [2, 4])
def test_slab_fftn_with_the_transpose_fused_into_the_store(emul, P):
    """The multi-GPU slab fftn of scirs_b200/distributed.py (SlabFftn, mode "p2p") with the ranks run one after the other on
    host arrays: axis 2 locally, axis 1 with the scatter store writing block q straight into rank q's receive buffer
    (peer memory over NVLink on the device), axis 0 on the received [n0][s1][n2] slab."""
    rng = np.random.default_rng(8)
    n0, n1, n2 = 16, 64, 32
    s0, s1 = n0 // P, n1 // P
    X = rng.standard_normal((n0, n1, n2)) + 1j * rng.standard_normal((n0, n1, n2))
    recv = [np.zeros((n0, s1, n2), dtype=np.complex128) for _ in range(P)]
    block = s0 * s1 * n2 * 16
    for r in range(P):
        x = np.ascontiguousarray(X[r * s0:(r + 1) * s0])
        work = np.zeros_like(x)
        rc, d = emul([s0, n1, n2], [2], x, work)
        assert rc == 0, d
        rc, d = emul.scatter([s0, n1, n2], [1], work, [recv[q].ctypes.data + r * block for q in range(P)], P)
        assert rc == 0, d
    ref = np.fft.fftn(X)
    for q in range(P):
        out = np.zeros_like(recv[q])
        rc, d = emul([n0, s1, n2], [0], recv[q], out)
        assert rc == 0, d
        assert rel(out, ref[:, q * s1:(q + 1) * s1, :]) < 1e-14


@pytest.mark.parametrize("P,shape", [(2, (16, 64, 32)), (4, (16, 64, 32)), (2, (32, 16, 64)), (8, (128, 64, 64)), (2, (16, 64, 32)),
                                     (4, (32, 128, 32)), (2, (64, 64, 64))])
@pytest.mark.parametrize("inverse", [False, True])
def test_slab_fftn_natural_layout_second_exchange(emul, P, shape, inverse):
    """csrc/dist.cu, layout SFC_SLAB_NATURAL, ranks run one after the other on host arrays: the axis-0 pass stores rows
    [q*s0, (q+1)*s0) straight into rank q's [s0][n1][n2] output at column offset r*s1 (scatter_pitch = n1*n2), which
    restores axis-0 slabs without a separate transpose — the same plans sfc_dist_plan_create builds."""
    rng = np.random.default_rng(9)
    n0, n1, n2 = shape
    s0, s1 = n0 // P, n1 // P
    X = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    recv = [np.zeros((n0, s1, n2), dtype=np.complex128) for _ in range(P)]
    block = s0 * s1 * n2 * 16
    for r in range(P):
        x = np.ascontiguousarray(X[r * s0:(r + 1) * s0])
        work = np.zeros_like(x)
        rc, d = emul([s0, n1, n2], [2], x, work, inverse=inverse)
        assert rc == 0, d
        rc, d = emul.scatter([s0, n1, n2], [1], work, [recv[q].ctypes.data + r * block for q in range(P)], P, inverse=inverse)
        assert rc == 0, d
    outs = [np.zeros((s0, n1, n2), dtype=np.complex128) for _ in range(P)]
    for r in range(P):
        rc, d = emul.scatter([n0, s1, n2], [0], recv[r], [outs[q].ctypes.data + r * s1 * n2 * 16 for q in range(P)], P,
                             pitch=n1 * n2, inverse=inverse, scale=0.5)
        assert rc == 0, d
    ref = (np.fft.ifftn(X) * X.size if inverse else np.fft.fftn(X)) * 0.5
    for q in range(P):
        assert rel(outs[q], ref[q * s0:(q + 1) * s0]) < 1e-14


@pytest.mark.parametrize("P,shape,nwin", [(2, (128, 128, 64), 2), (2, (256, 128, 128), 4), (4, (256, 256, 64), 4)])
@pytest.mark.parametrize("natural", [False, True])
def test_slab_fftn_column_windows(emul, P, shape, nwin, natural):
    """csrc/dist.cu, pipelined exchange: the axis-1 scatter pass and the axis-0 pass executed window by window over column blocks of
    n2 (Plan::exec with an ExecWindow, PassParams::win_*), on the same dense arrays — block j of pass C only needs block j of
    every peer's pass B, which is what lets the device overlap them.  Ranks and windows run one after the other here."""
    rng = np.random.default_rng(10)
    n0, n1, n2 = shape
    s0, s1 = n0 // P, n1 // P
    X = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    recv = [np.zeros((n0, s1, n2), dtype=np.complex128) for _ in range(P)]
    block = s0 * s1 * n2 * 16
    for r in range(P):
        x = np.ascontiguousarray(X[r * s0:(r + 1) * s0])
        work = np.zeros_like(x)
        rc, d = emul([s0, n1, n2], [2], x, work)
        assert rc == 0, d
        rc, d = emul.windows([s0, n1, n2], [1], work, None, [recv[q].ctypes.data + r * block for q in range(P)], P, n2, nwin)
        assert rc == 0, d
    ref = np.fft.fftn(X) * 0.5
    if not natural:
        for q in range(P):
            out = np.zeros_like(recv[q])
            rc, d = emul.windows([n0, s1, n2], [0], recv[q], out, [], 0, n2, nwin, scale=0.5)
            assert rc == 0, d
            assert rel(out, ref[:, q * s1:(q + 1) * s1, :]) < 1e-14
    else:
        outs = [np.zeros((s0, n1, n2), dtype=np.complex128) for _ in range(P)]
        for r in range(P):
            rc, d = emul.windows([n0, s1, n2], [0], recv[r], None, [outs[q].ctypes.data + r * s1 * n2 * 16 for q in range(P)], P, n2, nwin,
                                 pitch=n1 * n2, scale=0.5)
            assert rc == 0, d
        for q in range(P):
            assert rel(outs[q], ref[q * s0:(q + 1) * s0]) < 1e-14


@pytest.mark.parametrize("shape,axes", [([5, 9], [1]), ([4, 27], [1]), ([7, 81], [1]), ([3, 243], [1]), ([2, 729], [1]), ([2, 2187], [1]),
                                         ([81, 6], [0]), ([3, 243, 5], [1]), ([2, 6561], [1]), ([1, 19683], [1]), ([2, 59049], [1]),
                                         ([27, 9, 4], [0, 1]), ([1, 177147], [1])])
@pytest.mark.parametrize("inverse", [False, True])
def test_power_of_three_tiles(emul, shape, axes, inverse):
    """csrc/r3_tile.cuh through the planner: lengths 3^k as radix-9/3 Stockham tiles (one pass up to 2187, a two-pass four-step
    above: 3^13 = 729 x 2187 is BASELINE configs[3]) instead of the padded Bluestein convolution.  rustfft plans `Radix3` for
    these lengths."""
    rng = np.random.default_rng(33)
    x = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    y = np.zeros(shape, dtype=np.complex128)
    rc, d = emul(shape, axes, x, y, inverse=inverse, scale=0.25)
    assert rc == 0, d
    assert "power-of-three tile" in d, d
    ref = (np.fft.ifftn(x, axes=axes) * np.prod([shape[a] for a in axes]) if inverse else np.fft.fftn(x, axes=axes)) * 0.25
    assert rel(y, ref) < 2e-14, rel(y, ref)
