"""CPU: the oracle against the reference's known-answer tests and the committed golden vectors.

Mirrors scirs2-fft's own tests for the path: doctests of fft/algorithms.rs and
rfft.rs, unit tests rfft.rs:926-1032, planning.rs:733-754, src/bin/accuracy_comparison.rs:83-267
with its tolerances (:423-431).
"""
import os

import numpy as np
import pytest

from oracle import scirs2_fft_oracle as orc

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))
ENGINES = ["scipy", "c"]


@pytest.fixture(scope="module", autouse=True)
def _built(build_artifacts):
    return build_artifacts


@pytest.mark.parametrize("engine", ENGINES)
def test_doctest_dc_component(engine):
    # fft/algorithms.rs:117-130
    s = orc.fft([1.0, 2.0, 3.0, 4.0], None, engine)
    assert abs(s[0].real - 10.0) < 1e-10 and abs(s[0].imag) < 1e-10
    np.testing.assert_allclose(s, G["kat_1234_out"], atol=1e-12)


@pytest.mark.parametrize("engine", ENGINES)
def test_doctest_roundtrip(engine):
    # fft/algorithms.rs:191-209
    x = np.array([1.0, 2.0, 3.0, 4.0])
    r = orc.ifft(orc.fft(x, None, engine), None, engine)
    assert np.max(np.abs(r.real - x)) < 1e-10 and np.max(np.abs(r.imag)) < 1e-10


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("n", [64, 128, 256, 512, 1024])
def test_kat_pure_sine(engine, n):
    # accuracy_comparison.rs:83-121, tolerance 1e-10 (:423-431)
    s = orc.fft(G[f"kat_sine_{n}_in"], None, engine)
    assert np.max(np.abs(s - G[f"kat_sine_{n}_out"])) < 1e-10


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("n", [64, 256, 1024])
def test_kat_parseval_and_roundtrip(engine, n):
    # accuracy_comparison.rs:123-200
    x = G[f"kat_roundtrip_{n}_in"]
    s = orc.fft(x, None, engine)
    e_t, e_f = np.sum(np.abs(x) ** 2), np.sum(np.abs(s) ** 2) / n
    assert abs(e_t - e_f) / e_t < 1e-10
    assert np.max(np.abs(orc.ifft(s, None, engine) - x)) < 1e-10


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("n", [16, 32, 64])
def test_kat_fft2_sine(engine, n):
    # accuracy_comparison.rs:202-267, mean error < 1e-8
    s = orc.fft2(G[f"kat_sine2d_{n}_in"], None, None, None, engine)
    assert np.mean(np.abs(s - G[f"kat_sine2d_{n}_out"])) < 1e-8
    assert abs(orc.fft2(G["kat_2x2_in"], engine=engine)[0, 0].real - 10.0) < 1e-10  # algorithms.rs:280-292


@pytest.mark.parametrize("engine", ENGINES)
def test_kat_impulse_and_rfft(engine):
    # planning.rs:733-754 ; rfft.rs:926-1032
    s = orc.backend_fft(G["kat_impulse_in"], engine)
    assert np.max(np.abs(np.abs(s) - 1.0)) < 1e-10
    sig = np.array([1.0, 2.0, 3.0, 4.0])
    sp = orc.rfft(sig, None, engine)
    assert sp.shape == (3,) and abs(sp[0].real - 10.0) < 1e-10
    sp8 = orc.rfft(sig, 8, engine)
    assert sp8.shape == (5,) and abs(sp8[0].real - 10.0) < 1e-10
    rs = orc.rfft(G["kat_rsine16_in"], None, engine)
    assert abs(abs(rs[2].imag) - 8.0) < 1e-10
    np.testing.assert_allclose(rs, G["kat_rsine16_out"], atol=1e-12)
    # rfft2 keeps n_rows/2+1 ROWS (rfft.rs:226-229)
    assert orc.rfft2(np.arange(12.0).reshape(4, 3), engine=engine).shape == (3, 3)


@pytest.mark.parametrize("engine", ENGINES)
def test_doctest_fftn_roundtrip(engine):
    # fft/algorithms.rs:560-574, :725-755 (2x2x2 round trip to 1e-10)
    v = np.arange(8.0).reshape(2, 2, 2)
    r = orc.ifftn(orc.fftn(v, engine=engine), engine=engine)
    assert np.max(np.abs(r - v)) < 1e-10


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("n", [3, 5, 7, 12, 17, 100, 127, 243, 1000])
def test_golden_lengths(engine, n):
    x = G[f"ora_fft_n{n}_in"]
    assert orc.rel_l2(orc.fft(x, n, engine), G[f"ora_fft_n{n}_out"]) < 5e-15
    assert orc.rel_l2(orc.ifft(x, n, engine), G[f"ora_ifft_n{n}_out"]) < 5e-15


@pytest.mark.parametrize("engine", ENGINES)
def test_golden_padding_quirks(engine):
    # fft(x, None) pads to next pow2 (algorithms.rs:142); ifft truncates back (:258-260)
    y = orc.fft(G["ora_fft_pad_in"], None, engine)
    assert y.shape == (128,) and orc.rel_l2(y, G["ora_fft_pad_out"]) < 5e-15
    z = orc.ifft(G["ora_ifft_pad_in"], None, engine)
    assert z.shape == (100,) and orc.rel_l2(z, G["ora_ifft_pad_out"]) < 5e-15
    r = orc.rfft(G["ora_rfft_in"], None, engine)
    assert r.shape == (46,) and orc.rel_l2(r, G["ora_rfft_out"]) < 5e-15
    assert orc.rel_l2(orc.irfft(r, 90, engine), G["ora_irfft_out"]) < 5e-15


@pytest.mark.parametrize("engine", ENGINES)
@pytest.mark.parametrize("tag,axes", [("all", [0, 1, 2]), ("a20", [2, 0]), ("a1", [1])])
def test_golden_fftn_norm_table(engine, tag, axes):
    v = G["ora_fftn_in"]
    for norm in (None, "backward", "ortho", "forward", "nonsense"):
        key = "none" if norm in (None, "nonsense") else norm
        assert orc.rel_l2(orc.fftn(v, None, axes, norm, engine=engine), G[f"ora_fftn_{tag}_{key}"]) < 5e-15
    for norm in (None, "backward", "ortho", "forward"):
        key = "backward" if norm is None else norm
        assert orc.rel_l2(orc.ifftn(v, None, axes, norm, engine=engine), G[f"ora_ifftn_{tag}_{key}"]) < 5e-15


@pytest.mark.parametrize("engine", ENGINES)
def test_golden_fft2(engine):
    a = G["ora_fft2_in"]
    assert orc.rel_l2(orc.fft2(a, (12, 10), engine=engine), G["ora_fft2_shape_12x10"]) < 5e-15
    assert orc.rel_l2(orc.fft2(a, None, None, "ortho", engine), G["ora_fft2_ortho"]) < 5e-15
    assert orc.rel_l2(orc.ifft2(a, engine=engine), G["ora_ifft2_default"]) < 5e-15
    # axes are validated, then ignored (algorithms.rs:309-314)
    assert orc.rel_l2(orc.fft2(a, None, (1, 0), None, engine), orc.fft2(a, engine=engine)) == 0.0
    with pytest.raises(orc.OracleError) as e:
        orc.fft2(a, None, (0, 0))
    assert e.value.msg == "Invalid axes for 2D FFT"


def test_engines_agree_and_match_extended_precision():
    rng = np.random.default_rng(5)
    for n in (1, 2, 6, 31, 64, 97, 360, 1024, 4099):
        x = rng.standard_normal((2, n)) + 1j * rng.standard_normal((2, n))
        ref = orc.dft_longdouble(x)
        assert orc.rel_l2(orc._process(x, False, "c"), ref) < 2e-15
        assert orc.rel_l2(orc._process(x, False, "scipy"), ref) < 2e-15
    # sampled extended-precision bins at a BASELINE-sized prime length
    x = rng.standard_normal(100003) + 1j * rng.standard_normal(100003)
    bins = [0, 1, 50001, 100002]
    ref = orc.dft_longdouble(x, bins=bins)
    assert orc.rel_l2(orc._process(x, False, "c")[bins], ref) < 1e-14


def test_error_cases():
    with pytest.raises(orc.OracleError) as e:
        orc.fft([])
    assert e.value.variant == "ValueError" and e.value.msg == "Input cannot be empty"
    with pytest.raises(orc.OracleError) as e:
        orc.fftn(np.zeros((2, 2)), None, [2])
    assert e.value.msg == "Axis 2 out of bounds for array of dimension 2"
    with pytest.raises(orc.OracleError) as e:
        orc.fftn(np.zeros((2, 2)), [2, 2, 2])
    assert "same number of dimensions" in e.value.msg
    with pytest.raises(orc.OracleError) as e:
        orc.irfftn(np.zeros((2, 2)), None, [3])
    assert e.value.variant == "DimensionError"
    with pytest.raises(orc.OracleError) as e:
        orc.irfftn(np.zeros((2, 2, 2)), [4], [0, 1])
    assert e.value.variant == "DimensionError"


def test_rfftn_irfftn_semantics():
    rng = np.random.default_rng(9)
    v = rng.standard_normal((4, 6, 8))
    s = orc.rfftn(v)
    assert s.shape == (4, 6, 5)
    assert orc.rfftn(v, [4, 6, 8]).shape == (4, 6, 8)  # slicing only when shape is None (rfft.rs:508-511)
    assert orc.rfftn(v, None, [2, 0]).shape == (3, 6, 8)  # last LISTED axis is halved
    np.testing.assert_allclose(orc.irfftn(s, [4, 6, 8]), v, atol=1e-12)
    np.testing.assert_allclose(orc.irfftn(s), v, atol=1e-12)  # shape None -> 2*(5-1)
    # the literal Hermitian sweep equals the closed form "own value, else conj of reflection, else 0"
    x = rng.standard_normal((3, 4, 3)) + 1j * rng.standard_normal((3, 4, 3))
    full = orc.reconstruct_hermitian_symmetry(x, [5, 4, 6], [0, 2])
    for idx in np.ndindex(5, 4, 6):
        if all(i < s_ for i, s_ in zip(idx, x.shape)):
            exp = x[idx]
        else:
            r = list(idx)
            for t in (0, 2):
                n = [5, 4, 6][t]
                if idx[t] != 0 and not (n % 2 == 0 and idx[t] == n // 2):
                    r[t] = n - idx[t]
            exp = np.conj(x[tuple(r)]) if all(i < s_ for i, s_ in zip(r, x.shape)) else 0
        assert full[idx] == exp
