"""GPU: seeded randomised differential test of the drop-in API against the oracle — random ranks,
shapes (powers of two, smooth and prime extents), axes lists (subsets, permutations, duplicates),
pad/crop shapes, norm strings and input dtypes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-12
EXTENTS = [1, 2, 3, 4, 5, 6, 7, 8, 9, 12, 15, 16, 17, 24, 31, 32, 33, 48, 64, 100]
NORMS = [None, "backward", "ortho", "forward", "other"]


@pytest.fixture(scope="module")
def sb(build_artifacts):
    import scirs_b200 as m
    from scirs_b200 import _lib

    lib = _lib.load()
    assert lib.sfc_device_count() >= 1
    m.error.check(lib.sfc_init(0))
    return m


@pytest.fixture(scope="module")
def orc():
    from oracle import scirs2_fft_oracle as o

    return o


def rand_array(rng, shape, kind):
    if kind == "c128":
        return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    if kind == "c64":
        return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)
    if kind == "f32":
        return rng.standard_normal(shape).astype(np.float32)
    if kind == "i32":
        return rng.integers(-50, 50, size=shape).astype(np.int32)
    return rng.standard_normal(shape)


@pytest.mark.parametrize("seed", range(40))
def test_random_fftn_ifftn(sb, orc, seed):
    rng = np.random.default_rng(1000 + seed)
    nd = int(rng.integers(1, 5))
    shape = [int(rng.choice(EXTENTS[: 14 if nd > 2 else len(EXTENTS)])) for _ in range(nd)]
    kind = rng.choice(["f64", "c128", "f32", "c64", "i32"])
    x = rand_array(rng, shape, kind)
    k = int(rng.integers(0, nd + 2))
    axes = [int(a) for a in rng.integers(0, nd, size=k)] if rng.random() < 0.8 else None
    out_shape = None
    if rng.random() < 0.3:
        out_shape = [int(rng.choice(EXTENTS[:14])) for _ in range(nd)]
    norm = NORMS[int(rng.integers(0, len(NORMS)))]
    for fn_g, fn_o in ((sb.fftn, orc.fftn), (sb.ifftn, orc.ifftn)):
        got, ref = fn_g(x, out_shape, axes, norm), fn_o(x, out_shape, axes, norm)
        assert got.shape == ref.shape
        assert orc.rel_l2(got, ref) < TOL, (shape, kind, axes, out_shape, norm)


@pytest.mark.parametrize("seed", range(30))
def test_random_real_nd(sb, orc, seed):
    rng = np.random.default_rng(2000 + seed)
    nd = int(rng.integers(1, 4))
    shape = [int(rng.choice(EXTENTS[1:14])) for _ in range(nd)]
    x = rand_array(rng, shape, rng.choice(["f64", "f32"]))
    k = int(rng.integers(1, nd + 1))
    axes = [int(a) for a in rng.permutation(nd)[:k]] if rng.random() < 0.7 else None
    norm = NORMS[int(rng.integers(0, len(NORMS)))]
    got, ref = sb.rfftn(x, None, axes, norm), orc.rfftn(x, None, axes, norm)
    assert got.shape == ref.shape and orc.rel_l2(got, ref) < TOL, (shape, axes, norm)
    # irfftn of that spectrum: default shape, explicit original shape, and an odd/padded shape
    sp = ref
    for oshape in (None, list(x.shape), [s + int(rng.integers(0, 3)) for s in sp.shape]):
        g2, r2 = sb.irfftn(sp, oshape, axes, norm), orc.irfftn(sp, oshape, axes, norm)
        assert g2.shape == r2.shape
        assert orc.rel_l2(g2, r2) < TOL or np.linalg.norm(r2) < 1e-9, (shape, axes, oshape, norm)


@pytest.mark.parametrize("seed", range(30))
def test_random_1d_and_2d(sb, orc, seed):
    rng = np.random.default_rng(3000 + seed)
    n = int(rng.integers(1, 700))
    kind = rng.choice(["f64", "c128", "f32", "c64"])
    x = rand_array(rng, [n], kind)
    m = None if rng.random() < 0.4 else int(rng.integers(1, 900))
    assert orc.rel_l2(sb.fft(x, m), orc.fft(x, m)) < TOL, (n, m, kind)
    assert orc.rel_l2(sb.ifft(x, m), orc.ifft(x, m)) < TOL, (n, m, kind)
    if kind in ("f64", "f32"):
        assert orc.rel_l2(sb.rfft(x, m), orc.rfft(x, m)) < TOL, (n, m, kind)
    spec = rand_array(rng, [int(rng.integers(2, 300))], "c128")
    mo = None if rng.random() < 0.3 else int(rng.integers(1, 700))
    g, r = sb.irfft(spec, mo), orc.irfft(spec, mo)
    assert g.shape == r.shape and (orc.rel_l2(g, r) < TOL or np.linalg.norm(r) < 1e-9), (spec.size, mo)
    a = rand_array(rng, [int(rng.integers(1, 40)), int(rng.integers(1, 40))], kind)
    shp = None if rng.random() < 0.5 else (int(rng.integers(1, 48)), int(rng.integers(1, 48)))
    norm = NORMS[int(rng.integers(0, len(NORMS)))]
    assert orc.rel_l2(sb.fft2(a, shp, None, norm), orc.fft2(a, shp, None, norm)) < TOL
    assert orc.rel_l2(sb.ifft2(a, shp, None, norm), orc.ifft2(a, shp, None, norm)) < TOL
    assert orc.rel_l2(sb.rfft2(a, shp), orc.rfft2(a, shp)) < TOL
    ax = int(rng.integers(0, 2))
    assert orc.rel_l2(sb.fft_strided(a, ax), orc.fft_strided(a, ax)) < TOL
    assert orc.rel_l2(sb.ifft_strided(a, ax), orc.ifft_strided(a, ax)) < TOL
