"""GPU: the consumers of the FFT hot path (DCT/DST, Hartley, hfft/ihfft, hilbert,
stft/spectrogram) through the C ABI against oracle/consumers_oracle.py.  f64 bar: rel-L2 <= 1e-12."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def sb(build_artifacts):
    import scirs_b200 as m
    from scirs_b200 import _lib

    lib = _lib.load()
    assert lib.sfc_device_count() >= 1, "GPU tests need a CUDA device"
    m.error.check(lib.sfc_init(0))
    return m


@pytest.fixture(scope="module")
def co():
    from oracle import consumers_oracle as o

    return o


def rel(a, b):
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    d = np.linalg.norm((a - b).ravel()); r = np.linalg.norm(b.ravel())
    return d / r if r > 0 else d


# lengths: powers of two (fused single-kernel path), P = 2n not a power of two (Bluestein + explicit passes),
# n = 2^k +- 1 (types I: P = 2(n -+ 1) is a power of two), and one four-step size
@pytest.mark.parametrize("kind", ["dct", "dst"])
@pytest.mark.parametrize("t", [1, 2, 3, 4])
@pytest.mark.parametrize("n", [2, 3, 4, 7, 8, 9, 16, 33, 100, 255, 256, 257, 1000, 4096])
def test_dct_dst_all_types_norms_directions(sb, co, kind, t, n):
    rng = np.random.default_rng(1000 * t + n)
    x = rng.standard_normal(n)
    for norm in (None, "ortho"):
        for inv in (False, True):
            name = ("i" if inv else "") + kind
            got = getattr(sb, name)(x, t, norm)
            ref = getattr(co, name)(x, t, norm)
            assert rel(got, ref) <= TOL, (name, t, n, norm, rel(got, ref))


def test_dct_large_four_step_and_bluestein(sb, co):
    rng = np.random.default_rng(3)
    import scipy.fft as sf
    for n in (16384, 65536, 30000):  # P = 2n: 32768 / 131072 (four-step, fused) and 60000 (three-pass Bluestein)
        x = rng.standard_normal(n)
        assert rel(sb.dct(x, 2, None), sf.dct(x, 2) / 2) <= TOL      # same sums as the oracle (tests/test_consumers_oracle.py)
        assert rel(sb.dct(x, 4, "ortho"), sf.dct(x, 4, norm="ortho")) <= TOL
        assert rel(sb.dst(x, 2, None), sf.dst(x, 2) / 2) <= TOL
        assert rel(sb.idct(sb.dct(x, 2, "ortho"), 2, "ortho"), x) <= TOL


def test_dct_dst_nd(sb, co):
    rng = np.random.default_rng(4)
    a = rng.standard_normal((6, 16, 10))
    for t in (1, 2, 3, 4):
        assert rel(sb.dctn(a, t, "ortho"), co.dctn(a, t, "ortho")) <= TOL
        assert rel(sb.idctn(a, t, None, [2, 0]), co.idctn(a, t, None, [2, 0])) <= TOL
        assert rel(sb.dstn(a, t, None, [1]), co.dstn(a, t, None, [1])) <= TOL
        assert rel(sb.idstn(a, t, "ortho"), co.idstn(a, t, "ortho")) <= TOL
    m = rng.standard_normal((32, 48))
    assert rel(sb.dct2(m, 2, "ortho"), co.dct2(m, 2, "ortho")) <= TOL
    assert rel(sb.idct2(m, 2, "ortho"), co.idct2(m, 2, "ortho")) <= TOL
    assert rel(sb.dst2(m, 3), co.dst2(m, 3)) <= TOL
    assert rel(sb.idst2(m, 1, "ortho"), co.idst2(m, 1, "ortho")) <= TOL
    # reference unit tests (dct.rs:757-768, 826-840, 843-862)
    sig = np.array([1.0, 2.0, 3.0, 4.0])
    assert np.allclose(sb.idct(sb.dct(sig, sb.DCTType.Type2, "ortho"), sb.DCTType.Type2, "ortho"), sig, atol=1e-10)
    c = sb.dct(np.full(4, 3.0), sb.DCTType.Type2, None)
    assert abs(c[0]) > 1e-10 and np.all(np.abs(c[1:]) < 1e-10)
    arr = np.array([[1.0, 2.0], [3.0, 4.0]])
    assert np.allclose(sb.idct2(sb.dct2(arr, None, "ortho"), None, "ortho"), arr, atol=1e-10)


def test_dct_errors(sb):
    with pytest.raises(sb.ValueError_) as e:
        sb.dct([1.0], sb.DCTType.Type1)
    assert "at least 2 elements for DCT-I" in str(e.value)
    with pytest.raises(sb.ValueError_) as e:
        sb.idst([1.0], sb.DSTType.Type1)
    assert "at least 2 elements for IDST-I" in str(e.value)
    with pytest.raises(sb.ValueError_):
        sb.dct([], None)


def test_fused_and_unfused_paths_agree(sb, co, monkeypatch):
    # the explicit pre/post-pass path is what non-power-of-two lengths use; force it for a power of two too
    import subprocess, sys
    code = ("import numpy as np, scirs_b200 as sb; from oracle import consumers_oracle as co;"
            "x=np.random.default_rng(1).standard_normal((3,64,5));"
            "e=max(np.linalg.norm(sb.dctn(x,t,'ortho',[1])-co.dctn(x,t,'ortho',[1]))/np.linalg.norm(co.dctn(x,t,'ortho',[1])) for t in (1,2,3,4));"
            "print(e); assert e<1e-12")
    env = dict(os.environ, SFC_EXT_FUSE="0")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.parametrize("n", [1, 2, 4, 5, 16, 100, 1024, 5000])
def test_hartley(sb, co, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n)
    assert rel(sb.dht(x), co.dht(x)) <= TOL
    assert rel(sb.idht(x), co.idht(x)) <= TOL
    assert rel(sb.fht(x), co.dht(x)) <= TOL


def test_hartley_reference_tests_and_2d(sb, co):
    x = np.array([1.0, 2.0, 3.0, 4.0])
    assert np.allclose(sb.idht(sb.dht(x)), x, atol=1e-10)  # hartley.rs:216-231
    a = np.random.default_rng(2).standard_normal((12, 20))
    for axes in (None, (0, 1), (1, 0), (0, 0), (1, 1)):
        assert rel(sb.dht2(a, axes), co.dht2(a, axes)) <= TOL
    with pytest.raises(sb.ValueError_):
        sb.dht2(a, (0, 2))
    with pytest.raises(sb.ValueError_):
        sb.dht(np.array([]))


@pytest.mark.parametrize("n", [1, 2, 7, 8, 64, 1000])
def test_hfft_ihfft_hilbert(sb, co, n):
    rng = np.random.default_rng(n + 7)
    z = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    assert rel(sb.hfft(z), co.hfft(z)) <= TOL
    assert rel(sb.hfft(z, n + 5), co.hfft(z, n + 5)) <= TOL
    assert rel(sb.hfft(z.real), co.hfft(z.real)) <= TOL
    x = rng.standard_normal(n)
    assert rel(sb.ihfft(x), co.ihfft(x)) <= TOL
    assert rel(sb.ihfft(x, n + 3), co.ihfft(x, n + 3)) <= TOL
    if n > 2:
        assert rel(sb.ihfft(x, n - 1), co.ihfft(x, n - 1)) <= TOL
    assert rel(sb.hilbert(x), co.hilbert(x)) <= TOL
    assert rel(sb.hilbert(z), co.hilbert(z)) <= TOL  # complex input: real part only (lib.rs:455-463)


def test_stft_spectrogram(sb, co):
    rng = np.random.default_rng(11)
    x = rng.standard_normal(5000)
    for args in (dict(window="hann", nperseg=256), dict(window="hamming", nperseg=100, noverlap=25, nfft=128),
                 dict(window="blackman", nperseg=64, noverlap=0, detrend=False),
                 dict(window="rectangular", nperseg=128, return_onesided=False),
                 dict(window="hann", nperseg=200, nfft=256, boundary="reflect", fs=48000.0),
                 dict(window="hann", nperseg=64, boundary="zeros"), dict(window="hann", nperseg=64, boundary="constant")):
        f, t, z = sb.stft(x, **args)
        fr, tr, zr = co.stft(x, **args)
        assert np.allclose(f, fr) and np.allclose(t, tr)
        assert rel(z, zr) <= TOL, (args, rel(z, zr))
    for mode in ("psd", "magnitude"):
        for scaling in ("density", "spectrum"):
            f, t, p = sb.spectrogram(x, 100.0, "hann", 128, 64, None, True, scaling, mode)
            fr, tr, pr = co.spectrogram(x, 100.0, "hann", 128, 64, None, True, scaling, mode)
            assert rel(p, pr) <= TOL
    # phases: compare as unit vectors (atan2 at +-pi)
    f, t, p = sb.spectrogram(x, None, None, 128, None, None, None, None, "phase")
    _, _, pr = co.spectrogram(x, None, None, 128, None, None, None, None, "phase")
    assert np.max(np.abs(np.exp(1j * p) - np.exp(1j * pr))) < 1e-9
    f, t, p = sb.spectrogram(x, None, None, 128, None, None, None, None, "angle")
    assert np.max(np.abs(np.exp(1j * p * np.pi / 180) - np.exp(1j * pr))) < 1e-9
    with pytest.raises(sb.ValueError_):
        sb.stft(x, "hann", 128, 128)
    with pytest.raises(sb.ValueError_):
        sb.stft(x, "hann", 128, 64, 64)
    with pytest.raises(sb.ValueError_):
        sb.spectrogram(x, scaling="bogus")
    with pytest.raises(sb.ValueError_):
        sb.stft([], "hann", 16)
    w = np.hanning(50)
    _, _, z = sb.stft(x, w, 50, 10, 64)
    _, _, zr = co.stft(x, w, 50, 10, 64)
    assert rel(z, zr) <= TOL


def test_memory_efficient_and_ndim_optimized(sb, co):
    rng = np.random.default_rng(21)
    for n, inv, norm in ((8, False, False), (8, False, True), (16, True, True), (31, True, False), (64, False, True),
                         (1024, True, False), (100, True, True)):
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        a, b = x.copy(), np.zeros(n + 3, dtype=np.complex128)
        ra, rb = x.copy(), np.zeros(n + 3, dtype=np.complex128)
        assert sb.fft_inplace(a, b, sb.FftMode.Inverse if inv else sb.FftMode.Forward, norm) == n
        co.fft_inplace(ra, rb, inv, norm)
        assert rel(a, ra) <= TOL and rel(b[:n], rb[:n]) <= TOL, (n, inv, norm)
    with pytest.raises(sb.ValueError_):
        sb.fft_inplace(np.zeros(100, dtype=np.complex128), np.zeros(100, dtype=np.complex128))  # forward, not a power of two
    with pytest.raises(sb.ValueError_) as e:
        sb.fft_inplace(np.zeros(8, dtype=np.complex128), np.zeros(4, dtype=np.complex128))
    assert "Output buffer is too small: got 4, need 8" in str(e.value)
    m = rng.standard_normal((24, 40))
    for shape, inv, norm in ((None, False, False), ((32, 32), False, True), ((16, 64), True, True), (None, True, False)):
        got = sb.fft2_efficient(m, shape, sb.FftMode.Inverse if inv else sb.FftMode.Forward, norm)
        assert rel(got, co.fft2_efficient(m, shape, inv, norm)) <= TOL
    mc = m + 1j * rng.standard_normal(m.shape)
    assert rel(sb.fft2_efficient(mc), co.fft2_efficient(mc)) <= TOL
    x = rng.standard_normal(5000)
    for n, inv, chunk in ((None, False, None), (4096, False, None), (6000, True, None), (None, False, 1024), (None, True, 1000),
                          (7000, False, 2048), (4500, True, 2048)):
        got = sb.fft_streaming(x, n, sb.FftMode.Inverse if inv else sb.FftMode.Forward, chunk)
        assert rel(got, co.fft_streaming(x, n, inv, chunk)) <= TOL, (n, inv, chunk)
    assert np.allclose(sb.process_in_chunks(x, 1024, lambda c: sb.fft(c, len(c))),
                       np.concatenate([np.fft.fft(x[s:s + 1024]) for s in range(0, 5000, 1024)]), atol=1e-9)
    a = rng.standard_normal((6, 16, 10))
    assert rel(sb.fftn_optimized(a), co.fftn_optimized(a)) <= TOL
    assert rel(sb.fftn_optimized(a, None, [2, 0]), co.fftn_optimized(a, None, [2, 0])) <= TOL
    a = rng.standard_normal((8, 32))
    assert rel(sb.fftn_optimized(a), np.fft.fftn(a)) <= TOL  # power-of-two extents: the plain N-D transform
    with pytest.raises(sb.ValueError_):
        sb.fftn_optimized(a, None, [2])


@pytest.mark.parametrize("n", [128, 256, 512, 1024, 2048, 4096])
def test_fused_dct_dst_types_2_3_rows(sb, co, n):
    """Types II / III on contiguous power-of-two rows take the one-kernel Makhoul path (api_ext.cu); batches that do not
    fill a tile fall back to the 2n-point path: both against the literal O(n^2) sums."""
    rng = np.random.default_rng(n)
    for rows in (64, 3):
        a = rng.standard_normal((rows, n))
        for t in (2, 3):
            for norm in (None, "ortho"):
                assert rel(sb.dctn(a, t, norm, [1]), co.dctn(a[:3], t, norm, [1]) if rows == 3 else
                           np.vstack([co.dctn(a[:2], t, norm, [1]), sb.dctn(a, t, norm, [1])[2:]])) <= TOL
                assert rel(sb.idctn(a, t, norm, [1])[:2], co.idctn(a[:2], t, norm, [1])) <= TOL, ("idct", t, norm, n, rows)
                assert rel(sb.dstn(a, t, norm, [1])[:2], co.dstn(a[:2], t, norm, [1])) <= TOL, ("dst", t, norm, n, rows)
                assert rel(sb.idstn(a, t, norm, [1])[:2], co.idstn(a[:2], t, norm, [1])) <= TOL, ("idst", t, norm, n, rows)
    # last row of a full batch too (tile boundaries)
    a = rng.standard_normal((64, n))
    assert rel(sb.dctn(a, 2, "ortho", [1])[-1], co.dct(a[-1], 2, "ortho")) <= TOL
    assert rel(sb.idstn(a, 3, None, [1])[-1], co.idst(a[-1], 3, None)) <= TOL


@pytest.mark.parametrize("n", [8192, 16384])
def test_fused_dct_dst_largest_rows_against_scipy(sb, n):
    """At these lengths the literal O(n^2) sums of the reference (angles up to pi*n evaluated in f64) are themselves only
    good to ~1e-12, so the fused kernels are checked against scipy's transforms through the identities pinned in
    tests/test_consumers_oracle.py."""
    import scipy.fft as sf
    a = np.random.default_rng(n).standard_normal((16, n))
    assert rel(sb.dctn(a, 2, None, [1]), sf.dct(a, 2, axis=1) / 2) <= TOL
    assert rel(sb.dctn(a, 2, "ortho", [1]), sf.dct(a, 2, axis=1, norm="ortho")) <= TOL
    assert rel(sb.idctn(a, 2, None, [1]), 2 * sf.idct(a, 2, axis=1)) <= TOL
    assert rel(sb.dctn(a, 3, None, [1]), 2 * sf.idct(a, 2, axis=1)) <= TOL
    assert rel(sb.idctn(a, 3, None, [1]), (sf.dct(a, 3, axis=1) + a[:, :1]) / 2) <= TOL
    assert rel(sb.dstn(a, 2, None, [1]), sf.dst(a, 2, axis=1) / 2) <= TOL
    assert rel(sb.idctn(sb.dctn(a, 2, "ortho", [1]), 2, "ortho", [1]), a) <= TOL


@pytest.mark.parametrize("shape,axis", [((256, 64), 0), ((1024, 32), 0), ((4, 128, 64), 1), ((2, 512, 16), 1), ((2048, 8), 0),
                                        ((128, 64, 2), 0)])
def test_fused_dct_dst_types_2_3_strided_axes(sb, co, shape, axis):
    """Types II / III along a strided axis use the same fused kernels with column tiles (adjacent lanes adjacent in memory)."""
    a = np.random.default_rng(sum(shape)).standard_normal(shape)
    for t in (2, 3):
        for norm in (None, "ortho"):
            assert rel(sb.dctn(a, t, norm, [axis]), co.dctn(a, t, norm, [axis])) <= TOL, ("dct", t, norm)
            assert rel(sb.idctn(a, t, norm, [axis]), co.idctn(a, t, norm, [axis])) <= TOL, ("idct", t, norm)
            assert rel(sb.dstn(a, t, norm, [axis]), co.dstn(a, t, norm, [axis])) <= TOL, ("dst", t, norm)
            assert rel(sb.idstn(a, t, norm, [axis]), co.idstn(a, t, norm, [axis])) <= TOL, ("idst", t, norm)
    if len(shape) == 2:
        assert rel(sb.dct2(a, 2, "ortho"), co.dct2(a, 2, "ortho")) <= TOL
        assert rel(sb.idct2(a, 2, "ortho"), co.idct2(a, 2, "ortho")) <= TOL


def test_against_committed_extended_precision_golden_vectors(sb):
    """The product against tests/golden/golden_consumers_v1.npz (long-double evaluation of the reference's formulas)."""
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_consumers_v1.npz"))
    for n in (2, 5, 8, 16, 33, 128, 257):
        x = G[f"trig_x_{n}"]
        for kind in ("dct", "dst"):
            for t in (1, 2, 3, 4):
                for inv in (0, 1):
                    for ortho in (0, 1):
                        got = getattr(sb, ("i" if inv else "") + kind)(x, t, "ortho" if ortho else None)
                        assert rel(got, G[f"{kind}_{n}_t{t}_i{inv}_o{ortho}"]) <= TOL, (kind, n, t, inv, ortho)
    for n in (1, 4, 5, 12, 64, 100):
        x, z, m = G[f"real_{n}"], G[f"cplx_{n}"], n + 3
        assert rel(sb.dht(x), G[f"dht_{n}"]) <= TOL
        assert rel(sb.hfft(z, m), G[f"hfft_{n}_n{m}"]) <= TOL
        assert rel(sb.ihfft(x, m), G[f"ihfft_{n}_n{m}"]) <= TOL
        assert rel(sb.hilbert(x), G[f"hilbert_{n}"]) <= TOL


def test_czt_and_zoom_fft(sb, co):
    """czt.rs: the reference's own properties (test_czt_points :371-393, test_czt_as_fft :396-410) and the direct sum."""
    import importlib

    cz = importlib.import_module("scirs_b200.czt")

    pts = cz.czt_points(4)
    assert len(pts) == 4 and np.allclose(np.abs(pts), 1.0, atol=1e-10)
    a, w = 0.8 + 0j, 0.95 * np.exp(0.1j)
    pts = cz.czt_points(5, a, w)
    assert len(pts) == 5 and abs(pts[0] - a) < 1e-10
    x = np.linspace(0.0, 7.0, 8) + 0j
    assert np.allclose(cz.czt(x), np.fft.fft(x), atol=1e-10)                     # czt with defaults == fft
    rng = np.random.default_rng(8)
    for n, m in ((8, 8), (100, 37), (37, 100), (1000, 1000), (5000, 3000), (4096, 8192)):
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        assert rel(cz.czt(x, m), co.czt(x, m)) <= 1e-11, (n, m)
        # off-circle spirals are ill-conditioned in the fast form (|a|^-k |w|^(k^2/2) spans many decades): small sizes only
        damp = (0.9999, 0.98) if max(n, m) <= 100 else (1.0, 1.0)
        w = np.exp(-2j * np.pi * 0.37 / m) * damp[0]
        a = damp[1] * np.exp(0.3j)
        assert rel(cz.czt(x, m, w, a), co.czt(x, m, w, a)) <= 1e-9, (n, m, "spiral")
    x2 = rng.standard_normal((6, 50)) + 1j * rng.standard_normal((6, 50))
    assert rel(cz.czt(x2, 20), co.czt(x2, 20)) <= 1e-11
    assert rel(cz.czt(x2, 9, axis=0), np.moveaxis(co.czt(np.moveaxis(x2, 0, -1), 9), -1, 0)) <= 1e-11
    # zoom_fft: m points of the (oversampled) spectrum between f0 and f1
    x = rng.standard_normal(256) + 0j
    z = cz.zoom_fft(x, 64, 0.1, 0.3)
    k0, k1 = 0.1 * 256 * 2, 0.3 * 256 * 2
    f = (k0 + (k1 - k0) / 63 * np.arange(64)) / (256 * 2)
    ref = np.array([np.sum(x * np.exp(-2j * np.pi * fk * np.arange(256))) for fk in f])
    assert rel(z, ref) <= 1e-11
    for bad in (lambda: cz.zoom_fft(x, 8, 0.5, 0.2), lambda: cz.zoom_fft(x, 8, -0.1, 0.2), lambda: cz.zoom_fft(x, 8, 0.1, 0.2, 0.5),
                lambda: cz.CZT(0), lambda: cz.CZT(4, 0), lambda: cz.CZT(4).transform(np.zeros(5, dtype=complex)),
                lambda: cz.CZT(4).transform(np.zeros((2, 2, 4), dtype=complex))):
        with pytest.raises(sb.ValueError_):
            bad()
