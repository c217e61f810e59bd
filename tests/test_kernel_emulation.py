"""CPU: the tile kernel's SOURCE (scirs_b200/csrc/fft_tile.cuh) compiled for the host against a stand-in cuda_runtime.h
(tests/emul/), its threads run as OS threads with a barrier for __syncthreads().  This checks the index logic of kernel
flavours on a machine without a GPU — in particular the two flavours written after the round's GPU budget was spent
(TM_FAST_DCT4, TM_FAST_2D) — with the pass parameters set exactly as the planner sets them (plan.cu add_dct4 /
add_fft2_three_pass).  Test infrastructure only: nothing here is part of the product library, and it says nothing about
races, memory ordering or performance on the device.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL = os.path.join(ROOT, "tests", "emul")

MAP_ROW, MAP_COL = 0, 1
LD_C, ST_C, ST_TW = 0, 0, 1
F_IN_NOMASK, F_OUT_NOMASK, F_TRIG_SINE = 1 << 6, 1 << 7, 1 << 9


class IoDesc(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("batch_stride", C.c_int64), ("outer_stride", C.c_int64), ("inner_stride", C.c_int64),
                ("elem_stride", C.c_int64), ("len", C.c_int64), ("pos_es", C.c_int64), ("pos_ls", C.c_int64)]


class PassParams(C.Structure):  # pass_params.h, field for field
    _fields_ = [("in_", IoDesc), ("out", IoDesc), ("nlanes", C.c_uint32), ("inner_count", C.c_uint32),
                ("tiles_per_batch", C.c_uint32), ("nbatch_fast", C.c_uint32), ("total_tiles", C.c_uint32),
                ("tile_group_shift", C.c_uint32), ("map_in", C.c_int32), ("map_out", C.c_int32), ("ld_op", C.c_int32),
                ("st_op", C.c_int32), ("flags", C.c_uint32), ("tw", C.c_void_p), ("aux_in", C.c_void_p),
                ("aux_out", C.c_void_p), ("tw_lo", C.c_void_p), ("tw_hi", C.c_void_p), ("tw_shift", C.c_int32),
                ("mid", C.c_void_p), ("mid_es", C.c_int64), ("mid_ls", C.c_int64), ("mid_is", C.c_int64),
                ("ld_tw_lo", C.c_void_p), ("ld_tw_hi", C.c_void_p), ("ld_tw_shift", C.c_int32), ("rtw", C.c_void_p),
                ("chirp_lo", C.c_void_p), ("chirp_hi", C.c_void_p), ("chirp_shift", C.c_int32), ("chirp_mod", C.c_uint64),
                ("chirp_q_in", C.c_double * 2), ("chirp_q_out", C.c_double * 2), ("scale", C.c_double),
                ("scale_dc", C.c_double), ("peer_shift", C.c_int32), ("peer_out", C.c_void_p * 16),
                # alignas(64) CUtensorMap of the tensor-map tiles (never used by the host emulation), then the struct's tail padding
                ("_pad_tmap", C.c_uint8 * 8), ("tmap_in", C.c_uint8 * 128), ("tmap_box_rows", C.c_int32),
                ("tmap_split", C.c_int32), ("win_row_tiles", C.c_uint32), ("win_first", C.c_uint32), ("win_len", C.c_uint32),
                ("_pad_tail", C.c_uint8 * 44)]


@pytest.fixture(scope="module")
def emul():
    out = os.path.join(EMUL, "_build", "libtile_emul.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    srcs = [os.path.join(EMUL, "tile_emul.cpp"), os.path.join(EMUL, "cuda_runtime.h"),
            os.path.join(ROOT, "scirs_b200", "csrc", "fft_tile.cuh"), os.path.join(ROOT, "scirs_b200", "csrc", "pass_params.h")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-pthread", "-I" + EMUL,
                        "-I" + os.path.join(ROOT, "scirs_b200", "csrc"), "-I" + os.path.join(ROOT, "include"), "-o", out, srcs[0]],
                       check=True)
    lib = C.CDLL(out)
    lib.emul_run.argtypes = [C.c_longlong, C.POINTER(PassParams), C.c_uint]
    assert lib.emul_sizeof_params() == C.sizeof(PassParams), "tests/test_kernel_emulation.py is out of step with pass_params.h"
    return lib


def roots(n, count=None, mult=1):
    j = np.arange(n if count is None else count)
    return np.ascontiguousarray(np.exp(-2j * np.pi * (j * mult) / n))


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def io(d, arr, bs, os_, is_, es, ln, pes, pls):
    d.ptr, d.batch_stride, d.outer_stride, d.inner_stride, d.elem_stride, d.len, d.pos_es, d.pos_ls = ptr(arr), bs, os_, is_, es, ln, pes, pls


def run(lib, L, TL, mode, p, grid):
    assert lib.emul_run(L * 1000000 + TL * 1000 + mode, C.byref(p), grid) == 0, "kernel instantiation missing in tile_emul.cpp"


def test_emulator_reproduces_a_validated_flavour(emul):
    """Sanity of the emulator itself: TM_FAST_C2C rows and strided lanes (GPU-validated flavours) against numpy."""
    rng = np.random.default_rng(0)
    L, TL, rows = 512, 8, 16
    x = rng.standard_normal((rows, L)) + 1j * rng.standard_normal((rows, L))
    y = np.zeros_like(x)
    tw = roots(L)
    p = PassParams()
    io(p.in_, x, 0, L, 1, 1, L, 1, 0)
    io(p.out, y, 0, L, 1, 1, L, 1, 0)
    p.nlanes, p.inner_count, p.tiles_per_batch = rows, 1, rows // TL
    p.map_in = p.map_out = MAP_ROW
    p.ld_op, p.st_op, p.flags, p.scale, p.peer_shift = LD_C, ST_C, F_IN_NOMASK | F_OUT_NOMASK, 1.0, -1
    p.tw = ptr(tw)
    run(emul, L, TL, 1, p, rows // TL)
    assert np.linalg.norm(y - np.fft.fft(x, axis=1)) / np.linalg.norm(y) < 1e-14
    # strided lanes: transform axis 0 of [L][cols], adjacent lanes adjacent in memory
    cols = 16
    x = rng.standard_normal((L, cols)) + 1j * rng.standard_normal((L, cols))
    y = np.zeros_like(x)
    p = PassParams()
    io(p.in_, x, 0, L * cols, 1, cols, L, 1, 0)
    io(p.out, y, 0, L * cols, 1, cols, L, 1, 0)
    p.nlanes, p.inner_count, p.tiles_per_batch = cols, cols, cols // TL
    p.map_in = p.map_out = MAP_COL
    p.ld_op, p.st_op, p.flags, p.scale, p.peer_shift = LD_C, ST_C, F_IN_NOMASK | F_OUT_NOMASK, 0.5, -1
    p.tw = ptr(tw)
    run(emul, L, TL, 1, p, cols // TL)
    assert np.linalg.norm(y - 0.5 * np.fft.fft(x, axis=0)) / np.linalg.norm(y) < 1e-14


@pytest.mark.parametrize("L,TL,O,I", [(64, 32, 32, 1), (512, 4, 8, 1), (256, 16, 2, 16)])
@pytest.mark.parametrize("sine", [False, True])
def test_dct4_kernel_logic(emul, L, TL, O, I, sine):
    """TM_FAST_DCT4 with the parameters of PlanBuilder::add_dct4 (rows: I == 1; strided axis: I adjacent lanes)."""
    from oracle import consumers_oracle as co

    n = 2 * L
    rng = np.random.default_rng(L + I)
    x = rng.standard_normal((O, n, I))
    y = np.zeros_like(x)
    j = np.arange(L)
    pre = np.ascontiguousarray(np.exp(-1j * np.pi * (4 * j + 1) / (4 * n)))
    post = roots(2 * n, L)
    tw = roots(L)
    p = PassParams()
    io(p.in_, x, 0, n * I, 1, I, n, 1, 0)
    io(p.out, y, 0, n * I, 1, I, n, 1, 0)
    lanes = O * I
    p.nlanes, p.inner_count, p.tiles_per_batch = lanes, I, lanes // TL
    p.map_in = p.map_out = MAP_COL if I > 1 else MAP_ROW
    p.ld_op, p.st_op, p.scale, p.peer_shift = LD_C, ST_C, 1.0, -1
    p.flags = F_IN_NOMASK | F_OUT_NOMASK | (F_TRIG_SINE if sine else 0)
    p.tw, p.aux_in, p.aux_out = ptr(tw), ptr(pre), ptr(post)
    run(emul, L, TL, 7, p, lanes // TL)
    # the un-normalised type-IV sums: co.dct(.., 4, None) is the plain cosine sum, co.dst(.., 4, None) twice the sine sum
    ref = (co.dstn(x, 4, None, [1]) / 2.0) if sine else co.dctn(x, 4, None, [1])
    assert np.linalg.norm(y - ref) / np.linalg.norm(ref) < 1e-13


def test_fft2_three_pass_plan_logic(emul):
    """The three passes of PlanBuilder::add_fft2_three_pass on a scaled-down array: R = A*16 rows, C = 32*Cb columns with
    A = 256, Cb = 8 (the full-size plan has A = 512, Cb = 256; the kernels and every stride formula are the same)."""
    rng = np.random.default_rng(9)
    LA, LB, A, Cb = 16, 32, 256, 8
    R, Cn = A * LA, LB * Cb
    x = rng.standard_normal((R, Cn)) + 1j * rng.standard_normal((R, Cn))
    ms, sa, out = np.zeros_like(x), np.zeros_like(x), np.zeros_like(x)
    # pass 1: four-step pass A of axis 0 (L1 = A, L2 = 16), ST_TW with the two-level table of W_R
    lg = int(np.log2(R))
    sh = (lg + 1) // 2
    lo, hi = roots(R, 1 << sh), roots(R, max(R >> sh, 1), 1 << sh)
    p = PassParams()
    io(p.in_, x, R * Cn, Cn, 1, LA * Cn, R, LA, 1)
    io(p.out, ms, R * Cn, Cn, 1, LA * Cn, R, LA, 1)
    p.nlanes, p.inner_count = LA * Cn, Cn
    TL1 = 8
    p.tiles_per_batch = LA * Cn // TL1
    p.map_in = p.map_out = MAP_COL
    p.ld_op, p.st_op, p.flags, p.scale, p.peer_shift = LD_C, ST_TW, F_IN_NOMASK | F_OUT_NOMASK, 1.0, -1
    tw1 = roots(A)
    p.tw, p.tw_lo, p.tw_hi, p.tw_shift = ptr(tw1), ptr(lo), ptr(hi), sh
    run(emul, A, TL1, 1, p, LA * Cn // TL1)
    # pass 2: the 16 x 32 two-dimensional tile
    twc, tw2 = roots(Cn), roots(LA * LB)
    p = PassParams()
    io(p.in_, ms, 0, LA * Cn, 1, Cb, LA * LB, 1, 0)
    io(p.out, sa, 0, Cn, 1, Cb, LA * LB, 1, 0)
    p.mid_es, p.mid_ls = A * Cn, Cb
    p.nlanes, p.inner_count, p.tiles_per_batch = A * Cb, Cb, A * Cb // 8
    p.map_in = p.map_out = MAP_COL
    p.ld_op, p.st_op, p.flags, p.scale, p.peer_shift = LD_C, ST_C, F_IN_NOMASK | F_OUT_NOMASK, 1.0, -1
    p.tw, p.aux_out = ptr(tw2), ptr(twc)
    run(emul, LA * LB, 8, 8, p, A * Cb // 8)
    # pass 3: four-step pass B of axis 1 (L1 = 32, L2 = Cb): here Cb = 8 is below the tile sizes, so numpy stands in for it
    for kc1 in range(LB):
        out[:, kc1 + LB * np.arange(Cb)] = np.fft.fft(sa[:, kc1 * Cb:(kc1 + 1) * Cb], axis=1)
    ref = np.fft.fft2(x)
    assert np.linalg.norm(out - ref) / np.linalg.norm(ref) < 1e-13


def test_fft2_three_pass_plan_pass3_logic(emul):
    """Pass 3 of add_fft2_three_pass at full width (C = 8192 = 32 * 256) on a few rows: contiguous 256-point segments in,
    output column kc1 + 32*kc2 (four-step pass B of axis 1; one batch per row)."""
    rng = np.random.default_rng(10)
    LB, Cb, R = 32, 256, 3
    Cn = LB * Cb
    sa = rng.standard_normal((R, Cn)) + 1j * rng.standard_normal((R, Cn))
    out = np.zeros_like(sa)
    TL = 8
    p = PassParams()
    io(p.in_, sa, Cn, Cb, 1, 1, Cb, 1, 0)
    io(p.out, out, Cn, 1, 1, LB, Cn, LB, 1)
    p.nlanes, p.inner_count, p.tiles_per_batch = LB, 1, LB // TL
    p.map_in, p.map_out = MAP_ROW, MAP_COL
    p.ld_op, p.st_op, p.flags, p.scale, p.peer_shift = LD_C, ST_C, F_IN_NOMASK | F_OUT_NOMASK, 0.25, -1
    tw = roots(Cb)
    p.tw = ptr(tw)
    run(emul, Cb, TL, 1, p, (LB // TL) * R)
    ref = np.zeros_like(sa)
    for kc1 in range(LB):
        ref[:, kc1 + LB * np.arange(Cb)] = 0.25 * np.fft.fft(sa[:, kc1 * Cb:(kc1 + 1) * Cb], axis=1)
    assert np.linalg.norm(out - ref) / np.linalg.norm(ref) < 1e-14
