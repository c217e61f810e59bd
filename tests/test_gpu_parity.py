"""GPU: the CUDA path through the C ABI against the oracle and the committed golden vectors.

f64 bar: relative L2 <= 1e-12 (north_star); f32 bar: <= 1e-5 against the f64 oracle of the
same (f32-rounded) inputs.  Reads like the reference's own tests."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL64 = 1e-12
TOL32 = 1e-5
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))


@pytest.fixture(scope="module")
def sb(build_artifacts):
    import scirs_b200 as m
    from scirs_b200 import _lib

    lib = _lib.load()
    assert lib.sfc_device_count() >= 1, "GPU tests need a CUDA device"
    m.error.check(lib.sfc_init(0))
    return m


@pytest.fixture(scope="module")
def orc():
    from oracle import scirs2_fft_oracle as o

    return o


def cplx(rng, *s):
    return rng.standard_normal(s) + 1j * rng.standard_normal(s)


# ------------------------------------------------------------------ reference's own tests


def test_reference_doctests_and_unit_tests(sb):
    s = sb.fft([1.0, 2.0, 3.0, 4.0])  # algorithms.rs:117-130
    assert abs(s[0].real - 10.0) < 1e-10 and abs(s[0].imag) < 1e-10
    r = sb.ifft(s)  # :191-209
    assert np.max(np.abs(r.real - [1, 2, 3, 4])) < 1e-10 and np.max(np.abs(r.imag)) < 1e-10
    assert abs(sb.fft2(np.array([[1.0, 2.0], [3.0, 4.0]]))[0, 0].real - 10.0) < 1e-10  # :280-292
    v = np.arange(8.0).reshape(2, 2, 2)
    assert np.max(np.abs(sb.ifftn(sb.fftn(v)) - v)) < 1e-10  # :560-574, :725-755
    sig = np.array([1.0, 2.0, 3.0, 4.0])
    sp = sb.rfft(sig)  # rfft.rs:926-966
    assert sp.shape == (3,) and abs(sp[0].real - 10.0) < 1e-10
    np.testing.assert_allclose(sb.irfft(sp, 4), sig, atol=1e-10)
    sp8 = sb.rfft(sig, 8)
    assert sp8.shape == (5,) and abs(sp8[0].real - 10.0) < 1e-10
    rs = sb.rfft(G["kat_rsine16_in"])  # rfft.rs:995-1032
    assert abs(abs(rs[2].imag) - 8.0) < 1e-10
    assert sb.rfft2(np.arange(12.0).reshape(4, 3)).shape == (3, 3)  # rfft.rs:226-229
    assert np.max(np.abs(sb.fft_simd(sig) - sb.fft(sig))) < 1e-10  # tests/simd_fft_test.rs:6-29


@pytest.mark.parametrize("n", [64, 128, 256, 512, 1024])
def test_kat_pure_sine(sb, n):
    # accuracy_comparison.rs:83-121 (max err < 1e-10)
    assert np.max(np.abs(sb.fft(G[f"kat_sine_{n}_in"]) - G[f"kat_sine_{n}_out"])) < 1e-10


@pytest.mark.parametrize("n", [64, 256, 1024])
def test_kat_parseval_roundtrip(sb, n):
    x = G[f"kat_roundtrip_{n}_in"]
    s = sb.fft(x)
    assert abs(np.sum(np.abs(x) ** 2) - np.sum(np.abs(s) ** 2) / n) / np.sum(np.abs(x) ** 2) < 1e-10
    assert np.max(np.abs(sb.ifft(s) - x)) < 1e-10


@pytest.mark.parametrize("n", [16, 32, 64])
def test_kat_fft2_sine(sb, n):
    assert np.mean(np.abs(sb.fft2(G[f"kat_sine2d_{n}_in"]) - G[f"kat_sine2d_{n}_out"])) < 1e-8


def test_backend_trait(sb):
    b = sb.get_backend_manager().get_backend()
    assert b.name() == "cuda_fft" and b.is_available()
    out = np.empty(8, dtype=np.complex128)
    b.fft(G["kat_impulse_in"], out)  # planning.rs:733-754 / backend.rs:350-386
    assert np.max(np.abs(np.abs(out) - 1.0)) < 1e-10
    back = np.empty(8, dtype=np.complex128)
    b.ifft(out, back)  # 1/n normalised, backend.rs:149-152
    assert np.max(np.abs(back - G["kat_impulse_in"])) < 1e-12
    with pytest.raises(sb.ValueError_):
        b.fft_sized(out, back, 4)
    ex = sb.FftPlanExecutor([4, 2], True)
    o2 = np.empty(8, dtype=np.complex128)
    ex.execute(G["kat_impulse_in"], o2)
    assert np.max(np.abs(o2 - 1.0)) < 1e-12
    with pytest.raises(sb.ValueError_):
        ex.execute(np.zeros(7, dtype=np.complex128), o2)


# ------------------------------------------------------------------ golden vectors


@pytest.mark.parametrize("n", [3, 5, 7, 12, 17, 100, 127, 243, 1000])
def test_golden_lengths(sb, orc, n):
    x = G[f"ora_fft_n{n}_in"]
    assert orc.rel_l2(sb.fft(x, n), G[f"ora_fft_n{n}_out"]) < TOL64
    assert orc.rel_l2(sb.ifft(x, n), G[f"ora_ifft_n{n}_out"]) < TOL64


def test_golden_padding_quirks(sb, orc):
    y = sb.fft(G["ora_fft_pad_in"])
    assert y.shape == (128,) and orc.rel_l2(y, G["ora_fft_pad_out"]) < TOL64
    z = sb.ifft(G["ora_ifft_pad_in"])
    assert z.shape == (100,) and orc.rel_l2(z, G["ora_ifft_pad_out"]) < TOL64
    r = sb.rfft(G["ora_rfft_in"])
    assert r.shape == (46,) and orc.rel_l2(r, G["ora_rfft_out"]) < TOL64
    assert orc.rel_l2(sb.irfft(r, 90), G["ora_irfft_out"]) < TOL64


@pytest.mark.parametrize("tag,axes", [("all", [0, 1, 2]), ("a20", [2, 0]), ("a1", [1])])
def test_golden_fftn_norm_table(sb, orc, tag, axes):
    v = G["ora_fftn_in"]
    for norm in (None, "backward", "ortho", "forward", "nonsense"):
        key = "none" if norm in (None, "nonsense") else norm
        assert orc.rel_l2(sb.fftn(v, None, axes, norm), G[f"ora_fftn_{tag}_{key}"]) < TOL64
    for norm in (None, "backward", "ortho", "forward"):
        key = "backward" if norm is None else norm
        assert orc.rel_l2(sb.ifftn(v, None, axes, norm), G[f"ora_ifftn_{tag}_{key}"]) < TOL64


def test_golden_fft2(sb, orc):
    a = G["ora_fft2_in"]
    assert orc.rel_l2(sb.fft2(a, (12, 10)), G["ora_fft2_shape_12x10"]) < TOL64
    assert orc.rel_l2(sb.fft2(a, None, None, "ortho"), G["ora_fft2_ortho"]) < TOL64
    assert orc.rel_l2(sb.ifft2(a), G["ora_ifft2_default"]) < TOL64
    assert np.array_equal(sb.fft2(a, None, (1, 0)), sb.fft2(a))  # axes ignored after validation


# ------------------------------------------------------------------ oracle parity, seeded


@pytest.mark.parametrize("lg", list(range(1, 14)))
def test_pow2_lengths_batched(sb, orc, lg):
    """every tile-kernel instantiation, both precisions, both directions, rows and columns"""
    import scipy.fft as sf

    rng = np.random.default_rng(lg)
    n = 1 << lg
    b = max(3, min(96, (1 << 15) // n))
    x = cplx(rng, b, n)
    for fwd in (True, False):
        ref = sf.fft(x, axis=1) if fwd else sf.ifft(x, axis=1, norm="forward")
        assert orc.rel_l2(sb.FftPlan([b, n], [1], "c2c", "f64", fwd).execute(x), ref.ravel()) < TOL64
        x32 = x.astype(np.complex64)
        ref32 = sf.fft(x32.astype(np.complex128), axis=1) if fwd else sf.ifft(x32.astype(np.complex128), axis=1, norm="forward")
        assert orc.rel_l2(sb.FftPlan([b, n], [1], "c2c", "f32", fwd).execute(x32), ref32.ravel()) < TOL32
    inner = 12 if n <= 2048 else 3
    xc = cplx(rng, 2, n, inner)
    assert orc.rel_l2(sb.FftPlan([2, n, inner], [1], "c2c", "f64").execute(xc), sf.fft(xc, axis=1).ravel()) < TOL64


@pytest.mark.parametrize("n", [1 << 14, 1 << 16, 1 << 20, 1 << 22, 1 << 24])  # 2^22 and up: three-level form
def test_four_step(sb, orc, n):
    import scipy.fft as sf

    rng = np.random.default_rng(n)
    x = cplx(rng, 2, n)
    assert orc.rel_l2(sb.FftPlan([2, n], [1], "c2c", "f64").execute(x), sf.fft(x, axis=1).ravel()) < TOL64
    assert orc.rel_l2(sb.fft(x[0]), orc.fft(x[0])) < TOL64  # config 1 entry point
    assert orc.rel_l2(sb.ifft(x[1]), orc.ifft(x[1])) < TOL64


@pytest.mark.parametrize("n", [6, 31, 100, 1009, 4095, 4097, 6561, 10007, 65537])
def test_bluestein(sb, orc, n):
    rng = np.random.default_rng(n)
    x = cplx(rng, n)
    assert orc.rel_l2(sb.fft(x, n), orc.fft(x, n)) < TOL64
    assert orc.rel_l2(sb.ifft(x, n), orc.ifft(x, n)) < TOL64
    xr = rng.standard_normal(n)
    assert orc.rel_l2(sb.rfft(xr), orc.rfft(xr)) < TOL64
    assert orc.rel_l2(sb.irfft(orc.rfft(xr), n), orc.irfft(orc.rfft(xr), n)) < TOL64


def test_bluestein_baseline_lengths_sampled_bins(sb, orc):
    """config 4 lengths (prime 1,000,003 and 3^13) against extended-precision sampled bins"""
    rng = np.random.default_rng(4)
    for n in (1000003, 1594323):
        x = cplx(rng, n)
        y = sb.fft(x, n)
        bins = [0, 1, 2, n // 3, n // 2, n - 1]
        ref = orc.dft_longdouble(x, bins=bins)
        assert orc.rel_l2(y[bins], ref) < TOL64
        assert orc.rel_l2(sb.ifft(y, n), x) < TOL64  # round trip over the whole vector


@pytest.mark.parametrize("n", [64, 256, 4096, 16384])
def test_real_fast_paths(sb, orc, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal((9, n))
    ref = np.stack([orc.rfft(r) for r in x])
    got = sb.rfft_batch(x)
    assert orc.rel_l2(got, ref) < TOL64
    assert orc.rel_l2(sb.irfft_batch(ref, n), np.stack([orc.irfft(r, n) for r in ref])) < TOL64
    x32 = x.astype(np.float32)
    ref32 = np.stack([orc.rfft(r) for r in x32])  # reference widens f32 to f64 (SURVEY 8c f32 note)
    got32 = sb.rfft_batch(x32)
    assert got32.dtype == np.complex64 and orc.rel_l2(got32, ref32) < TOL32
    assert orc.rel_l2(sb.irfft_batch(got32, n), x32) < TOL32
    # imag of DC / Nyquist never reaches the real output (reference takes Re of a complex ifft)
    dirty = ref.copy()
    dirty[:, 0] += 3j
    dirty[:, -1] -= 2j
    assert orc.rel_l2(sb.irfft_batch(dirty, n), np.stack([orc.irfft(r, n) for r in dirty])) < TOL64


def test_irfft_shape_cases(sb, orc):
    rng = np.random.default_rng(11)
    s = cplx(rng, 257)
    for n in (None, 512, 511, 100, 257, 258, 1200, 3):
        assert orc.rel_l2(sb.irfft(s, n), orc.irfft(s, n)) < TOL64, n
    # the reference's hard-coded return for (len 3, n 4) is NOT reproduced: we return the transform
    s3 = np.array([10.0, -2 + 2j, -2.0])
    np.testing.assert_allclose(sb.irfft(s3, 4), [1, 2, 3, 4], atol=1e-12)
    s3b = np.array([1.0, 5j, 7.0])
    assert orc.rel_l2(sb.irfft(s3b, 4), orc.irfft(s3b, 4)) < TOL64


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.complex128, np.int32])
def test_input_dtypes_widen_to_f64(sb, orc, dtype):
    rng = np.random.default_rng(3)
    x = (rng.standard_normal(300) * 100).astype(dtype) if not np.issubdtype(dtype, np.complexfloating) else cplx(rng, 300).astype(dtype)
    assert orc.rel_l2(sb.fft(x), orc.fft(x)) < TOL64
    assert orc.rel_l2(sb.fft(x, 200), orc.fft(x, 200)) < TOL64
    assert orc.rel_l2(sb.fftn(x.reshape(15, 20)), orc.fftn(x.reshape(15, 20))) < TOL64


def test_fft2_family(sb, orc):
    rng = np.random.default_rng(21)
    a = rng.standard_normal((24, 40))
    c = cplx(rng, 64, 128)
    for norm in (None, "backward", "ortho", "forward", "bogus"):
        assert orc.rel_l2(sb.fft2(a, None, None, norm), orc.fft2(a, None, None, norm)) < TOL64
        assert orc.rel_l2(sb.ifft2(c, None, None, norm), orc.ifft2(c, None, None, norm)) < TOL64
        assert orc.rel_l2(sb.fft2_parallel(c, None, None, norm, 4), orc.fft2(c, None, None, norm)) < TOL64
    for shp in ((32, 32), (16, 50), (30, 7)):
        assert orc.rel_l2(sb.fft2(a, shp), orc.fft2(a, shp)) < TOL64
        assert orc.rel_l2(sb.ifft2(a, shp), orc.ifft2(a, shp)) < TOL64
    assert orc.rel_l2(sb.rfft2(a), orc.rfft2(a)) < TOL64
    assert orc.rel_l2(sb.rfft2(a, (16, 16)), orc.rfft2(a, (16, 16))) < TOL64
    sp = orc.rfft2(a)
    assert orc.rel_l2(sb.irfft2(sp), orc.irfft2(sp)) < TOL64
    assert orc.rel_l2(sb.irfft2(sp, (24, 40)), orc.irfft2(sp, (24, 40))) < TOL64
    with pytest.raises(sb.ValueError_) as e:
        sb.fft2(a, None, (0, 0))
    assert str(e.value) == "Invalid axes for 2D FFT"
    with pytest.raises(sb.ValueError_) as e:
        sb.ifft2(a, None, (0, 2))
    assert str(e.value) == "Invalid axes for 2D IFFT"


def test_fftn_family(sb, orc):
    rng = np.random.default_rng(22)
    v = rng.standard_normal((6, 8, 10))
    w = cplx(rng, 4, 16, 3, 8)
    for axes in (None, [0], [1], [2], [2, 0], [0, 1, 2], [1, 1], []):
        for norm in (None, "ortho", "backward", "forward"):
            assert orc.rel_l2(sb.fftn(v, None, axes, norm), orc.fftn(v, None, axes, norm)) < TOL64, (axes, norm)
            assert orc.rel_l2(sb.ifftn(v, None, axes, norm), orc.ifftn(v, None, axes, norm)) < TOL64, (axes, norm)
    assert orc.rel_l2(sb.fftn(w, None, [3, 1]), orc.fftn(w, None, [3, 1])) < TOL64
    assert orc.rel_l2(sb.fftn(v, [8, 8, 8]), orc.fftn(v, [8, 8, 8])) < TOL64
    assert orc.rel_l2(sb.fftn(v, [4, 9, 16], [1, 2]), orc.fftn(v, [4, 9, 16], [1, 2])) < TOL64
    assert orc.rel_l2(sb.fft_strided(v, 1), orc.fft_strided(v, 1)) < TOL64
    assert orc.rel_l2(sb.fft_strided_complex(w, 2), orc.fft_strided(w, 2)) < TOL64
    assert orc.rel_l2(sb.ifft_strided(sb.fft_strided_complex(w, 1), 1), w) < 1e-10  # strided_fft.rs:284-299
    for call, msg in ((lambda: sb.fftn(v, None, [3]), "Axis 3 out of bounds for array of dimension 3"),
                      (lambda: sb.fftn(v, [2, 2]), "Output shape must have the same number of dimensions as input"),
                      (lambda: sb.fft_strided(v, 5), "Axis 5 is out of bounds for array with 3 dimensions")):
        with pytest.raises(sb.ValueError_) as e:
            call()
        assert str(e.value) == msg
    with pytest.raises(sb.ValueError_) as e:
        sb.fft(np.zeros(0))
    assert str(e.value) == "Input cannot be empty"


def test_rfftn_irfftn_family(sb, orc):
    rng = np.random.default_rng(23)
    v = rng.standard_normal((6, 8, 10))
    for axes in (None, [2], [0, 1], [2, 0], [1, 1]):
        for norm in (None, "ortho", "forward"):
            assert orc.rel_l2(sb.rfftn(v, None, axes, norm), orc.rfftn(v, None, axes, norm)) < TOL64, (axes, norm)
    assert sb.rfftn(v, [6, 8, 16]).shape == (6, 8, 16)
    assert orc.rel_l2(sb.rfftn(v, [6, 8, 16]), orc.rfftn(v, [6, 8, 16])) < TOL64
    sp = orc.rfftn(v)
    for shape, axes in ((None, None), ([6, 8, 10], None), (None, [2]), ([7, 9, 11], None), ([12], [2]),
                        ([6, 8, 10], [0, 2]), (None, [0, 1]), ([6, 8, 64], None)):
        got, ref = sb.irfftn(sp, shape, axes), orc.irfftn(sp, shape, axes)
        assert got.shape == ref.shape and orc.rel_l2(got, ref) < TOL64, (shape, axes)
    big = rng.standard_normal((4, 32, 128))
    assert orc.rel_l2(sb.irfftn(orc.rfftn(big)), big) < TOL64  # fused C2R path
    with pytest.raises(sb.DimensionError):
        sb.irfftn(sp, None, [5])
    with pytest.raises(sb.DimensionError):
        sb.irfftn(sp, [4], [0, 1])


def test_plan_cache_semantics(sb):
    """plan_cache.rs:241-287"""
    c = sb.get_global_cache()
    c.configure(128, 3600.0)
    x = np.ones(64, dtype=np.complex128)
    sb.fft(x)
    sb.fft(x)
    s = c.get_stats()
    assert (s.hit_count, s.miss_count, s.size) == (1, 1, 1) and s.hit_rate == 0.5
    c.configure(2, 3600.0)  # cap 2 -> size 2 after 3 inserts
    for n in (8, 16, 32):
        sb.fft(np.ones(n, dtype=np.complex128))
    assert c.get_stats().size == 2
    c.configure(128, 3600.0)
    c.set_enabled(False)  # disabled -> counters untouched
    sb.fft(x)
    sb.fft(x)
    s = c.get_stats()
    assert (s.hit_count, s.miss_count, s.size) == (0, 0, 0)
    c.set_enabled(True)


# ------------------------------------------------------------------ BASELINE sizes: size-independent properties


def _torch():
    import torch

    return torch


def test_full_size_batched_rfft_roundtrip_and_sampled_rows(sb, orc):
    """config 2: 65,536 x 4096 f64 — sampled rows against the oracle, whole-array round trip"""
    torch = _torch()
    b, n = 65536, 4096
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(b, n, dtype=torch.float64, device="cuda", generator=g)
    spec = torch.empty(b, n // 2 + 1, dtype=torch.complex128, device="cuda")
    back = torch.empty_like(x)
    s = torch.cuda.current_stream().cuda_stream
    sb.FftPlan([b, n], [1], "r2c", "f64").execute_device(x, spec, s)
    sb.FftPlan([b, n], [1], "c2r", "f64", scale=1.0 / n).execute_device(spec, back, s)
    torch.cuda.synchronize()
    assert float((back - x).norm() / x.norm()) < TOL64
    rows = [0, 1, 4095, 32768, 65535]
    ref = np.stack([orc.rfft(x[r].cpu().numpy()) for r in rows])
    assert orc.rel_l2(spec[rows].cpu().numpy(), ref) < TOL64
    # Parseval on the half spectrum
    e_t = float((x[rows] ** 2).sum())
    sp = spec[rows]
    e_f = float((sp.abs() ** 2).sum() * 2 - (sp[:, 0].abs() ** 2).sum() - (sp[:, -1].abs() ** 2).sum()) / n
    assert abs(e_t - e_f) / e_t < 1e-12


def test_full_size_fft2_linearity_and_impulse(sb, orc):
    """config 3: 8192 x 8192 c128 — impulse -> plane wave, linearity, inverse round trip"""
    torch = _torch()
    n = 8192
    s = torch.cuda.current_stream().cuda_stream
    fwd = sb.FftPlan([n, n], [1, 0], "c2c", "f64", True)
    inv = sb.FftPlan([n, n], [1, 0], "c2c", "f64", False, 1.0 / (n * n))
    a = torch.zeros(n, n, dtype=torch.complex128, device="cuda")
    a[3, 5] = 1.0
    out = torch.empty_like(a)
    fwd.execute_device(a, out, s)
    torch.cuda.synchronize()
    k0 = torch.arange(n, device="cuda", dtype=torch.float64)
    ph = -2 * np.pi * ((3 * k0[:, None] + 5 * k0[None, :]) % n) / n
    expect = torch.complex(torch.cos(ph), torch.sin(ph))
    assert float((out - expect).abs().max()) < 1e-12
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(n, n, 2, dtype=torch.float64, device="cuda", generator=g)
    xc = torch.view_as_complex(x)
    y = torch.empty_like(xc)
    fwd.execute_device(xc, y, s)
    z = torch.empty_like(xc)
    inv.execute_device(y, z, s)
    torch.cuda.synchronize()
    assert float((z - xc).norm() / xc.norm()) < TOL64
    # sampled output bins by direct summation of one row/column decomposition
    row = xc[:, :].cpu().numpy()
    col_fft = orc.dft_longdouble(row.T, bins=[7])[:, 0]  # sum over axis 0 at k0 = 7 for every column
    ref = orc.dft_longdouble(col_fft, bins=[0, 11, 8191])
    got = y[7, [0, 11, 8191]].cpu().numpy()
    assert orc.rel_l2(got, ref) < TOL64


def test_full_size_fftn_512_roundtrip_and_axis_checks(sb, orc):
    """config 5 (single GPU): 512^3 c128 — round trip + one lane per axis against the oracle"""
    torch = _torch()
    n = 512
    s = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda").manual_seed(6)
    x = torch.view_as_complex(torch.randn(n, n, n, 2, dtype=torch.float64, device="cuda", generator=g))
    y = torch.empty_like(x)
    for ax in (0, 1, 2):
        sb.FftPlan([n, n, n], [ax], "c2c", "f64").execute_device(x, y, s)
        torch.cuda.synchronize()
        idx = [slice(None) if d == ax else 17 * (d + 1) for d in range(3)]
        lane = x[tuple(idx)].cpu().numpy()
        assert orc.rel_l2(y[tuple(idx)].cpu().numpy(), orc.backend_fft(lane)) < TOL64
    sb.FftPlan([n, n, n], [0, 1, 2], "c2c", "f64").execute_device(x, y, s)
    z = torch.empty_like(x)
    sb.FftPlan([n, n, n], [0, 1, 2], "c2c", "f64", False, 1.0 / n ** 3).execute_device(y, z, s)
    torch.cuda.synchronize()
    assert float((z - x).norm() / x.norm()) < TOL64
    assert abs(complex(y[0, 0, 0]) - complex(x.sum())) / abs(complex(x.sum())) < 1e-10  # DC = sum


def test_full_size_fft_2pow20(sb, orc):
    """config 1: 2^20 c128 through the drop-in call, against the oracle"""
    rng = np.random.default_rng(1)
    x = cplx(rng, 1 << 20)
    y = sb.fft(x)
    assert orc.rel_l2(y, orc.fft(x)) < TOL64
    assert orc.rel_l2(y[[0, 1, 12345]], orc.dft_longdouble(x, bins=[0, 1, 12345])) < TOL64


def test_planner_and_parallel_executor_mirrors(sb, orc):
    """planning.rs:701-823, planning_parallel.rs:408-477"""
    rng = np.random.default_rng(31)
    ex = sb.ParallelExecutor([256], True)
    ins = [cplx(rng, 256) for _ in range(5)]
    outs = [np.empty(256, dtype=np.complex128) for _ in range(5)]
    times = ex.execute_batch(ins, outs)
    assert len(times) == 5
    for a, b in zip(ins, outs):
        assert orc.rel_l2(b, orc.backend_fft(a)) < TOL64
    with pytest.raises(sb.ValueError_) as e:
        ex.execute_batch(ins, outs[:4])
    assert str(e.value) == "Input and output counts must match"
    with pytest.raises(sb.ValueError_) as e:
        ex.execute_batch([np.zeros(8, dtype=np.complex128)], [np.zeros(256, dtype=np.complex128)])
    assert str(e.value) == "Input 0 has wrong size: expected 256, got 8"
    plan = sb.PlanBuilder().shape([8]).forward(True).backend(sb.PlannerBackend.CUDA).build()
    out = np.empty(8, dtype=np.complex128)
    plan.execute(G["kat_impulse_in"], out)  # planning.rs:733-754
    assert np.max(np.abs(np.abs(out) - 1.0)) < 1e-10
    sb.plan_ahead_of_time([64, 128])
    assert sb.with_backend("cuda_fft", lambda: sb.get_backend_manager().get_backend().name()) == "cuda_fft"
    assert orc.rel_l2(sb.without_cache(lambda: sb.fft(ins[0])), orc.fft(ins[0])) < TOL64


def test_full_size_bluestein_batch_256(sb, orc):
    """config 4: 256 signals x N = 1,000,003 (prime) and x 3^13, device resident — round trip over the
    whole batch, sampled extended-precision bins of two signals, linearity across the batch"""
    torch = _torch()
    s = torch.cuda.current_stream().cuda_stream
    for n, seed in ((1000003, 4), (1594323, 5)):
        b = 256
        g = torch.Generator(device="cuda").manual_seed(seed)
        x = torch.view_as_complex(torch.randn(b, n, 2, dtype=torch.float64, device="cuda", generator=g))
        y = torch.empty_like(x)
        sb.FftPlan([b, n], [1], "c2c", "f64", True).execute_device(x, y, s)
        torch.cuda.synchronize()
        bins = [0, 1, n // 2, n - 1]
        for row in (0, 255):
            ref = orc.dft_longdouble(x[row].cpu().numpy(), bins=bins)
            assert orc.rel_l2(y[row, bins].cpu().numpy(), ref) < TOL64
        z = torch.empty_like(x)
        sb.FftPlan([b, n], [1], "c2c", "f64", False, 1.0 / n).execute_device(y, z, s)
        torch.cuda.synchronize()
        assert float((z - x).norm() / x.norm()) < TOL64
        del x, y, z
        torch.cuda.empty_cache()


def test_full_size_fftn_1024_roundtrip(sb, orc):
    """config 5b on one GPU: 1024^3 c128 (17.2 GB) forward + inverse, DC bin = sum, one lane per axis"""
    torch = _torch()
    n = 1024
    s = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.view_as_complex(torch.randn(n, n, n, 2, dtype=torch.float64, device="cuda", generator=g))
    y = torch.empty_like(x)
    sb.FftPlan([n, n, n], [0, 1, 2], "c2c", "f64").execute_device(x, y, s)
    torch.cuda.synchronize()
    tot = complex(x.sum())
    assert abs(complex(y[0, 0, 0]) - tot) / abs(tot) < 1e-9
    # bin (1, 0, 0): sum_i0 w^i0 * (sum over the other two axes)
    plane = x.sum(dim=(1, 2)).cpu().numpy()
    ref = orc.dft_longdouble(plane, bins=[1, 513])
    assert orc.rel_l2(np.array([complex(y[1, 0, 0]), complex(y[513, 0, 0])]), ref) < 1e-10
    sb.FftPlan([n, n, n], [0, 1, 2], "c2c", "f64", False, 1.0 / n ** 3).execute_device(y, y, s)  # in place
    torch.cuda.synchronize()
    assert float((y - x).norm() / x.norm()) < TOL64


def test_power_of_three_lengths(sb, orc):
    """Lengths 3^k run on radix-9/3 tiles (csrc/r3_tile.cuh; rustfft plans Radix3 for them): one pass up to 3^7, a
    two-pass four-step up to 3^14.  Against the oracle (extended-precision DFT bins for the long ones), both directions, f32."""
    rng = np.random.default_rng(314)
    for k in range(2, 12):
        n = 3 ** k
        rows = 5 if k < 10 else 2
        a = cplx(rng, rows, n)
        p = sb.FftPlan([rows, n], [1])
        assert "radix-9/3" in p.describe(), p.describe()
        got = p.execute(a).reshape(rows, n)
        ref = np.stack([orc.fft(r, n) for r in a]) if n <= 6561 else np.fft.fft(a, axis=1)
        assert orc.rel_l2(got, ref) < TOL64, (n, orc.rel_l2(got, ref))
        inv = sb.FftPlan([rows, n], [1], "c2c", "f64", False, 1.0 / n).execute(got).reshape(rows, n)
        assert orc.rel_l2(inv, a) < TOL64
        got32 = sb.FftPlan([rows, n], [1], "c2c", "f32").execute(a.astype(np.complex64)).reshape(rows, n)
        assert orc.rel_l2(got32.astype(np.complex128), np.fft.fft(a.astype(np.complex64).astype(np.complex128), axis=1)) < TOL32
    # the free functions reach them too: fft(x, Some(3^9)), strided axes of an N-D array, partial last tiles
    v = rng.standard_normal(3 ** 9)
    assert orc.rel_l2(sb.fft(v, 3 ** 9), orc.fft(v, 3 ** 9)) < TOL64
    vol = cplx(rng, 27, 81, 10)
    assert orc.rel_l2(sb.fftn(vol, None, [1, 0]), orc.fftn(vol, None, [1, 0])) < TOL64
    assert orc.rel_l2(sb.ifftn(vol, None, [0]), orc.ifftn(vol, None, [0])) < TOL64


def test_full_size_3pow13_batch(sb, orc):
    """BASELINE configs[3]: 3^13 = 729 x 2187 as a two-pass transform, batch 32 here (256 in bench.py): round trip, Parseval and
    extended-precision bins of one row."""
    n, rows = 3 ** 13, 32
    rng = np.random.default_rng(13)
    a = cplx(rng, rows, n)
    p = sb.FftPlan([rows, n], [1])
    d = p.describe()
    assert "radix-9/3" in d and p.info["num_passes"] == 2, d
    got = p.execute(a).reshape(rows, n)
    back = sb.FftPlan([rows, n], [1], "c2c", "f64", False, 1.0 / n).execute(got).reshape(rows, n)
    assert orc.rel_l2(back, a) < TOL64
    assert abs(np.sum(np.abs(got) ** 2) / n / np.sum(np.abs(a) ** 2) - 1.0) < 1e-12
    bins = np.array([0, 1, 728, 729, 2187, n // 2, n - 1])
    j = np.arange(n, dtype=np.int64)
    ref = np.array([np.sum(a[3].astype(np.clongdouble) * np.exp(-2j * np.pi * ((j * int(b)) % n).astype(np.longdouble) / n)) for b in bins])
    assert np.max(np.abs(got[3][bins] - ref.astype(np.complex128))) / np.max(np.abs(ref)) < 1e-12
