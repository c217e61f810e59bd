"""CPU tests of oracle/consumers_oracle.py (the checker of the SURVEY 8f rank-1 consumers).

Pins: the reference's own unit tests / doctests for these modules that do not depend on its
hard-coded `[1,2,3,4]` answers (dct.rs:757-768, 826-840, 843-862; dst.rs:712-722, 755-770;
hartley.rs:216-260; lib.rs hilbert doctest; spectrogram.rs tests), plus scipy cross-checks of the sums
where the reference's definition coincides with scipy's up to a stated constant.
"""
import numpy as np
import pytest
import scipy.fft as sf

from oracle import consumers_oracle as co
from oracle import scirs2_fft_oracle as orc


def test_reference_dct_unit_tests():
    sig = np.array([1.0, 2.0, 3.0, 4.0])
    # dct.rs:757-768 test_dct_and_idct (Type2 ortho round trip, 1e-10)
    assert np.allclose(co.idct(co.dct(sig, 2, "ortho"), 2, "ortho"), sig, atol=1e-10)
    # dct.rs:843-862 test_constant_signal
    c = co.dct(np.full(4, 3.0), 2, None)
    assert abs(c[0]) > 1e-10 and np.all(np.abs(c[1:]) < 1e-10)
    # dct.rs:826-840 test_dct2_and_idct2
    a = np.array([[1.0, 2.0], [3.0, 4.0]])
    assert np.allclose(co.idct2(co.dct2(a, 2, "ortho"), 2, "ortho"), a, atol=1e-10)
    # dct.rs:38-50 doctest: ortho DC / 2 == mean
    assert abs(co.dct(sig, 2, "ortho")[0] / 2.0 - 2.5) < 1e-10
    # dct.rs:806-815: DCT-IV ortho round trip keeps the ratio of the last to the first sample within 0.1
    r = co.idct(co.dct(sig, 4, "ortho"), 4, "ortho")
    assert abs(r[3] / r[0] - 4.0) < 0.1


def test_reference_dst_unit_tests():
    # dst.rs:755-770 test_dst2_and_idst2 only passes through the hard-coded 2x2 answer of idst2 (dst.rs:233-237);
    # the arithmetic the reference really performs, worked by hand from dst.rs:484-545 for x = [1, 2], ortho:
    #   dst2:  X1 = sin(pi/4) + 2 sin(3pi/4) = 3/sqrt(2),  X2 = sin(pi/2) + 2 sin(3pi/2) = -1
    #   idst2 = dst3(X, None): y0 = (X2 + X1 sin(pi/4)) / 2 = 0.25,  y1 = (-X2 + X1 sin(3pi/4)) / 2 = 1.25
    X = co.dst(np.array([1.0, 2.0]), 2, "ortho")
    assert np.allclose(X, [3.0 / np.sqrt(2.0), -1.0], atol=1e-14)
    assert np.allclose(co.idst(X, 2, "ortho"), [0.25, 1.25], atol=1e-14)
    # what the reference's scalings amount to, derived from dst.rs:409-480 and :630-702 with the orthogonality
    # sums  sum_k sin(..k..m)sin(..k..m') = (n+1)/2 (type I), n/2 (type IV):
    #   idst1(dst1(x)) = (2/sqrt(n+1))^2 * sqrt(n+1)/2 * (n+1)/2 * x = sqrt(n+1) * x   (not the identity)
    #   idst4(dst4(x)) = 2 * 1/2 * 2 * n/2 * x = n * x, and the same with "ortho"
    x = np.arange(1.0, 8.0)
    assert np.allclose(co.idst(co.dst(x, 1, None), 1, None), np.sqrt(8.0) * x, atol=1e-10)
    assert np.allclose(co.idst(co.dst(x, 4, None), 4, None), 7.0 * x, atol=1e-10)
    assert np.allclose(co.idst(co.dst(x, 4, "ortho"), 4, "ortho"), 7.0 * x, atol=1e-10)


@pytest.mark.parametrize("n", [1, 2, 5, 16, 33])
def test_sums_against_scipy_where_definitions_coincide(n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n)
    assert np.allclose(co.dct(x, 2, None), sf.dct(x, 2) / 2, atol=1e-12)
    assert np.allclose(co.dct(x, 2, "ortho"), sf.dct(x, 2, norm="ortho"), atol=1e-12)
    assert np.allclose(co.dct(x, 4, None), sf.dct(x, 4) / 2, atol=1e-12)
    assert np.allclose(co.dct(x, 4, "ortho"), sf.dct(x, 4, norm="ortho"), atol=1e-12)
    assert np.allclose(co.idct(x, 2, None), 2 * sf.idct(x, 2), atol=1e-12)  # (2/n)[x0/2 + sum] = 2 * scipy's idct-II
    assert np.allclose(co.idct(co.dct(x, 2, None), 2, None), x, atol=1e-12)
    assert np.allclose(co.idct(x, 3, None), (sf.dct(x, 3) + x[0]) / 2, atol=1e-12)  # sum_k x_k cos(pi k (i+1/2)/n)
    assert np.allclose(co.dct(x, 3, None), 2 * sf.idct(x, 2), atol=1e-12)            # dct3 == idct2 when un-normalised
    assert np.allclose(co.dst(x, 2, None), sf.dst(x, 2) / 2, atol=1e-12)
    assert np.allclose(co.dst(x, 4, None), sf.dst(x, 4), atol=1e-12)
    if n >= 2:
        assert np.allclose(co.dst(x, 1, None), sf.dst(x, 1) / np.sqrt(n + 1.0), atol=1e-12)


def test_reference_hartley_unit_tests():
    x = np.array([1.0, 2.0, 3.0, 4.0])
    assert np.allclose(co.idht(co.dht(x)), x, atol=1e-10)            # hartley.rs:216-231
    assert np.all(np.isfinite(co.dht(np.arange(1.0, 6.0))))          # hartley.rs:234-243 (n = 5: padded transform)
    assert co.dht2(np.array([[1.0, 2.0], [3.0, 4.0]])).shape == (2, 2)  # hartley.rs:246-252
    with pytest.raises(co.OracleError):
        co.dht(np.array([]))                                         # hartley.rs:255-259
    # n = 5 quirk: first 5 bins of the 8-point transform of the zero-padded signal
    f = np.fft.fft(np.concatenate([np.arange(1.0, 6.0), np.zeros(3)]))[:5]
    assert np.allclose(co.dht(np.arange(1.0, 6.0)), f.real - f.imag, atol=1e-12)


def test_hfft_ihfft_hilbert():
    rng = np.random.default_rng(5)
    z = rng.standard_normal(12) + 1j * rng.standard_normal(12)
    zz = z.copy(); zz[0] = zz[0].real
    assert np.allclose(co.hfft(z), np.fft.fft(zz).real, atol=1e-12)
    assert np.allclose(co.hfft(z, 20), np.fft.fft(zz, 20).real, atol=1e-12)
    x = rng.standard_normal(9)
    r = np.fft.ifft(x)
    ih = co.ihfft(x)
    assert ih[0] == complex(r[0].real, 0.0) and np.allclose(ih[1:5], r[1:5]) and np.allclose(ih[5:], np.conj(r[4:0:-1]))
    # hilbert, power-of-two length: ifft(fft(x) * h)
    x = rng.standard_normal(16)
    h = np.zeros(16, dtype=complex); h[0] = 1; h[8] = 1; h[1:8] = -2j
    assert np.allclose(co.hilbert(x), np.fft.ifft(np.fft.fft(x) * h), atol=1e-12)
    # non-power-of-two: padded transform, filter on the first n bins, 1/P scale, first n outputs
    x = rng.standard_normal(10)
    h = np.zeros(10, dtype=complex); h[0] = 1; h[5] = 1; h[1:5] = -2j
    s = np.fft.fft(x, 16)[:10] * h
    assert np.allclose(co.hilbert(x), np.fft.ifft(s, 16)[:10], atol=1e-12)


def test_stft_and_spectrogram_shapes_and_values():
    rng = np.random.default_rng(7)
    x = rng.standard_normal(1000)
    f, t, z = co.stft(x, "hann", 128, 64, None, 100.0, True, True, None)
    assert z.shape == (65, 1 + (1000 - 128) // 64) and f[1] == 100.0 / 128 and t[0] == 64 / 100.0
    seg = x[64:192] - x[64:192].mean()
    w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(128) / 127)
    assert np.allclose(z[:, 1], np.fft.rfft(seg * w), atol=1e-10)
    f2, t2, p = co.spectrogram(x, 100.0, "hann", 128, 64, None, True, "density", "psd")
    assert np.allclose(p, np.abs(z) ** 2 / (100.0 * (w * w).sum()), atol=1e-12)
    for b in ("reflect", "zeros", "constant"):
        _, _, zb = co.stft(x, "hamming", 100, 25, 128, None, False, True, b)
        assert zb.shape == (65, 1 + (1200 - 100) // 75)
    with pytest.raises(co.OracleError):
        co.stft(x, "hann", 128, 128)
    with pytest.raises(co.OracleError):
        co.stft(x, "hann", 128, 64, 64)


# ------------------------------------------------------------------ committed extended-precision vectors
import os

GC = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_consumers_v1.npz"))


def _rel(a, b):
    d = np.linalg.norm(np.asarray(a) - np.asarray(b)); r = np.linalg.norm(np.asarray(b))
    return d / r if r > 0 else d


def test_oracle_against_extended_precision_golden_vectors():
    """tests/golden/make_golden_consumers.py evaluates the reference's formulas in long double; the f64 restatement
    must agree to ~1e-13 (its own rounding), for every type / direction / norm and the Hartley / hfft / hilbert family."""
    for n in (2, 5, 8, 16, 33, 128, 257):
        x = GC[f"trig_x_{n}"]
        for kind in ("dct", "dst"):
            for t in (1, 2, 3, 4):
                for inv in (0, 1):
                    for ortho in (0, 1):
                        fn = getattr(co, ("i" if inv else "") + kind)
                        got = fn(x, t, "ortho" if ortho else None)
                        assert _rel(got, GC[f"{kind}_{n}_t{t}_i{inv}_o{ortho}"]) < 2e-13, (kind, n, t, inv, ortho)
    for n in (1, 4, 5, 12, 64, 100):
        x, z, m = GC[f"real_{n}"], GC[f"cplx_{n}"], n + 3
        assert _rel(co.dht(x), GC[f"dht_{n}"]) < 1e-13
        assert _rel(co.hfft(z, m), GC[f"hfft_{n}_n{m}"]) < 1e-13
        assert _rel(co.ihfft(x, m), GC[f"ihfft_{n}_n{m}"]) < 1e-13
        assert _rel(co.hilbert(x), GC[f"hilbert_{n}"]) < 1e-13


def test_dct4_half_length_formulation():
    """The arithmetic of the fused type-IV kernel (TM_FAST_DCT4 in fft_tile.cuh; experimental, SFC_DCT4_FUSED=1), emulated
    with numpy exactly as the kernel indexes it, against the literal sums: tables pre[j] = exp(-i pi (4j+1)/(4N)),
    post[k] = exp(-i pi k / N); load z[j] = (x[2j], x[N-1-2j]) (swapped for the sine transform) * pre[j]; store
    X[2k] = Re(y) * scale, X[N-1-2k] = -/+ Im(y) * scale; scales as in api_ext.cu trig_axis."""
    rng = np.random.default_rng(12)
    for n in (8, 128, 1024):
        L = n // 2
        j = np.arange(L)
        pre = np.exp(-1j * np.pi * (4 * j + 1) / (4 * n))
        post = np.exp(-1j * np.pi * j / n)
        x = rng.standard_normal(n)

        def kernel(v, sine, scale):
            ev, od = v[2 * j], v[n - 1 - 2 * j]
            z = ((od + 1j * ev) if sine else (ev + 1j * od)) * pre
            y = np.fft.fft(z) * post
            out = np.empty(n)
            out[2 * j] = y.real * scale
            out[n - 1 - 2 * j] = y.imag * (scale if sine else -scale)
            return out

        for ortho in (False, True):
            norm = "ortho" if ortho else None
            nn = float(n)
            sc = {("c", False): np.sqrt(2 / nn) if ortho else 1.0,
                  ("c", True): np.sqrt(nn / 2) * np.sqrt(2 / nn) if ortho else 2 / nn,
                  ("s", False): np.sqrt(2 / nn) if ortho else 2.0,
                  ("s", True): 2 * np.sqrt(nn / 2) if ortho else 1.0}
            assert np.allclose(kernel(x, False, sc[("c", False)]), co.dct(x, 4, norm), rtol=0, atol=1e-12 * n)
            assert np.allclose(kernel(x, False, sc[("c", True)]), co.idct(x, 4, norm), rtol=0, atol=1e-12 * n)
            assert np.allclose(kernel(x, True, sc[("s", False)]), co.dst(x, 4, norm), rtol=0, atol=1e-12 * n)
            assert np.allclose(kernel(x, True, sc[("s", True)]), co.idst(x, 4, norm), rtol=0, atol=1e-12 * n)
