"""CPU tests for the scirs2-signal callers.

1. oracle/signal_oracle.py against the reference's own unit-test assertions (spectral.rs:743-935,
   streaming_stft.rs tests) and against scipy.signal where the two definitions agree up to the
   reference's extra 1/nperseg (spectral.rs:203-207 divides by fs * len on top of 1 / sum(w^2)).
2. the HOST logic of scirs_b200/signal.py (framing, detrending, bookkeeping, bin passes) with its four
   device transforms replaced by the oracle's — tests only; the product itself has no CPU transform.
"""
import numpy as np
import pytest
import scipy.signal as ss

from oracle import scirs2_fft_oracle as orc
from oracle import signal_oracle as so

import _signal_cases as sc


def test_reference_spectral_unit_tests():
    fs = 100.0
    x = np.sin(2 * np.pi * 10.0 * np.arange(1000) / fs)
    f, p = so.periodogram(x, fs)  # spectral.rs:744-770
    assert abs(f[np.argmax(p)] - 10.0) <= 1.0 and 500 <= len(f) <= 502
    xn = np.sin(2 * np.pi * 10.0 * np.arange(2000) / fs) + np.random.default_rng(0).uniform(-0.1, 0.1, 2000)
    f, p = so.welch(xn, fs, None, 256, 128)  # spectral.rs:773-802
    assert abs(f[np.argmax(p)] - 10.0) <= 1.0
    t = np.arange(2000) / 1000.0
    f, tt, Z = so.stft(np.sin(2 * np.pi * (10.0 + 50.0 * t) * t), 1000.0, None, 128, 64)  # spectral.rs:805-863
    assert Z.shape == (len(tt), len(f))
    assert f[np.argmax(np.abs(Z[-1]))] > f[np.argmax(np.abs(Z[0]))]
    for mode in ("psd", "magnitude", "phase"):  # spectral.rs:866-935
        _, _, S = so.spectrogram(x, fs, None, 128, None, None, None, None, mode)
        assert S.size > 0
        assert np.all(S >= 0.0) if mode != "phase" else np.all((S >= -np.pi) & (S <= np.pi))


def test_oracle_against_scipy_for_power_of_two_lengths():
    rng = np.random.default_rng(3)
    x = rng.standard_normal(1024)
    f, p = so.periodogram(x, 4.0, "hann")
    fr, pr = ss.periodogram(x, 4.0, window=np.array(so.get_window("hann", 1024)), detrend="constant",
                            return_onesided=False, scaling="density")
    assert np.allclose(f, fr[:512]) and np.allclose(p, pr[:512] / 1024, rtol=1e-10, atol=1e-18)
    x = rng.standard_normal(4000)
    f, p = so.welch(x, 2.0, "hamming", 256, 100)
    fr, pr = ss.welch(x, 2.0, window=np.array(so.get_window("hamming", 256)), nperseg=256, noverlap=100,
                      detrend="constant", return_onesided=False, scaling="density")
    assert np.allclose(f, fr[:128]) and np.allclose(p, pr[:128] / 256, rtol=1e-10, atol=1e-18)
    # linear detrend is scipy's
    seg = rng.standard_normal(100) + 0.3 * np.arange(100)
    assert np.allclose(so.apply_detrend(list(seg), "linear"), ss.detrend(seg, type="linear"), atol=1e-10)
    # periodic windows = scipy's fftbins windows
    for name in ("hann", "hamming", "blackman", "bartlett"):
        assert np.allclose(so.signal_window(name, 37, True), ss.get_window(name, 37, fftbins=True), atol=1e-14)
        assert np.allclose(so.signal_window(name, 38, False), ss.get_window(name, 38, fftbins=False), atol=1e-14)


def test_reference_streaming_unit_tests():
    s = so.StreamingStft(256, 128, center=False)  # streaming_stft.rs test_streaming_stft_processing
    r = s.process_frame(np.sin(2 * np.pi * 100.0 * np.arange(256) / 1000.0))
    assert r is not None and len(r) == 129
    s = so.StreamingStft(128, 64, center=False, magnitude_only=True)  # test_streaming_stft_magnitude
    assert len(s.process_frame([1.0] * 128)) == 65
    s = so.StreamingStft(128, 64, center=False)  # test_streaming_stft_batch_processing / _flush
    assert len(s.process_batch([1.0] * 512, 64)) > 0
    s = so.StreamingStft(128, 64, center=False)
    s.process_frame([1.0] * 100)
    assert len(s.flush()) > 0 or len(s.buf) == 0


def test_reference_cqt_unit_tests():
    q = 1.0 / (2.0 ** (1.0 / 12) - 1.0)
    kern, freqs, n_fft = so.cqt_kernel(55.0, 440.0, 12, q, 22050.0)  # cqt.rs:800-818 test_cqt_kernel
    assert len(kern) == int(np.ceil(np.log2(440.0 / 55.0) * 12)) and abs(freqs[0] - 55.0) < 0.1
    assert freqs[-1] >= 440.0 * 0.9 and n_fft & (n_fft - 1) == 0
    # cqt.rs test_create_window: symmetric hann, zero at both ends, ~1 in the middle
    w = so.signal_window("hann", 128, False)
    assert w[0] < 1e-10 and w[127] < 1e-10 and w[64] > 0.9
    # cqt.rs test_constant_q_transform (smaller rate so the literal oracle stays fast): the strongest bin is
    # within one bin of the tone
    fs = 4000.0
    x = np.sin(2 * np.pi * 440.0 * np.linspace(0.0, 0.5, 2000))
    kern, freqs, n_fft = so.cqt_kernel(110.0, 1000.0, 12, q, fs)
    c = so.cqt_frame(x, kern, n_fft)
    assert abs(int(np.argmax(np.abs(c))) - int(np.argmin(np.abs(freqs - 440.0)))) <= 1
    # cqt.rs test_cqt_spectrogram: shapes
    S, times = so.cqt_spectrogram(x, kern, n_fft, fs, 512)
    assert S.shape == (len(kern), int(np.ceil(2000 / 512.0))) and len(times) == S.shape[1]
    # cqt.rs test_chromagram: A (index 9 from C) carries energy, frames sum to one
    kern, freqs, n_fft = so.cqt_kernel(220.0, 880.0, 12, q, fs)
    ch = so.chromagram(so.cqt_frame(x, kern, n_fft).reshape(-1, 1), freqs)
    assert ch.shape == (12, 1) and ch[9, 0] > 0.1 and abs(ch[:, 0].sum() - 1.0) < 1e-6


def test_reference_hilbert_and_wvd_unit_tests():
    # hilbert.rs test_hilbert_transform: |analytic| of a 5 Hz cosine (n = 1000, fs = 100) is 1 within 0.1 in the middle half
    x = np.cos(2 * np.pi * 5.0 * np.arange(1000) / 100.0)
    a = so.hilbert(x)
    assert np.all(np.abs(np.abs(a[250:750]) - 1.0) < 0.1)
    # power-of-two length: the quirks vanish and the result is -i times (scipy's analytic signal minus DC / Nyquist terms);
    # checked through the magnitude of a bin-centred tone
    y = np.cos(2 * np.pi * 8 * np.arange(256) / 256.0)
    assert np.allclose(np.abs(so.hilbert(y)), 1.0, atol=1e-12)
    # wvd.rs test_wigner_ville_chirp / test_cross_wigner_ville: shapes, finite values, positive energy
    n = 64
    t = np.arange(n) / 64.0
    chirp = np.sin(2 * np.pi * (5.0 + 10.0 * t) * t)
    w = so.wigner_ville(chirp)
    assert w.shape == (n + 1, n) and np.all(np.isfinite(w))
    xw = so.cross_wvd(so.hilbert(chirp), so.hilbert(chirp[::-1].copy()), False)
    assert xw.shape == (n // 2 + 1, n) and np.all(np.isfinite(xw)) and np.sum(np.abs(xw)) > 0.0


@pytest.fixture()
def host_signal(monkeypatch):
    """scirs_b200.signal with its device transforms swapped for the oracle's (host-logic check only)."""
    import scirs_b200.signal as sg

    def spectra(x, nperseg, step, count, win, detrend, nfft, n_half, psd_scale=None):
        # what sfc_signal_spectra does on the device, with the oracle's own detrend
        P = 1
        while P < nfft:
            P *= 2
        rows = np.zeros((count, P))
        for i in range(count):
            rows[i, :nperseg] = np.array(so.apply_detrend(list(x[i * step:i * step + nperseg]), detrend)) * win
        Z = np.fft.rfft(rows, axis=1)[:, :n_half]
        return Z if psd_scale is None else (np.abs(Z) ** 2).sum(axis=0) * psd_scale

    monkeypatch.setattr(sg, "_segment_spectra", spectra)
    monkeypatch.setattr(sg, "fft", lambda x, n=None: orc.fft(np.asarray(x), n))
    monkeypatch.setattr(sg, "ifft", lambda x, n=None: orc.ifft(np.asarray(x), n))
    monkeypatch.setattr(sg, "rfft_batch", lambda m: np.fft.rfft(np.asarray(m, dtype=np.float64), axis=1))
    monkeypatch.setattr(sg, "fftn", lambda x, shape=None, axes=None: np.fft.fft(np.asarray(x), axis=axes[0]))
    monkeypatch.setattr(sg, "ifftn", lambda x, shape=None, axes=None: np.fft.ifft(np.asarray(x), axis=axes[0]))
    return sg


@pytest.mark.parametrize("name,fn", sc.cases(), ids=[c[0] for c in sc.cases()])
def test_host_logic_matches_oracle(host_signal, name, fn):
    got, ref = fn(host_signal, so)
    sc.compare(got, ref, 1e-11)


def test_host_errors_and_bookkeeping(host_signal):
    sg = host_signal
    from scirs_b200.error import ValueError_

    with pytest.raises(ValueError_):
        sg.periodogram([])
    with pytest.raises(ValueError_):
        sg.periodogram(np.ones(8), nfft=4)
    with pytest.raises(ValueError_):
        sg.welch(np.ones(64), nperseg=16, noverlap=16)
    with pytest.raises(ValueError_):
        sg.welch(np.ones(64), fs=-1.0)
    with pytest.raises(ValueError_):
        sg.stft(np.ones(64), nperseg=16, nfft=8)
    with pytest.raises(ValueError_):
        sg.stft(np.ones(64), window="nope")
    with pytest.raises(ValueError_):
        sg.spectrogram(np.ones(64), mode="complex")
    with pytest.raises(ValueError_):
        sg.spectral_subtraction(np.ones(20))
    with pytest.raises(ValueError_):
        sg.StreamingStft(sg.StreamingStftConfig(frame_length=8, hop_length=9))
    with pytest.raises(ValueError_):
        sg.compute_bispectrum(np.ones(3))
    # streaming_stft.rs test_streaming_stft_latency / test_real_time_stft
    st = sg.StreamingStft(sg.StreamingStftConfig(frame_length=512, hop_length=256, center=True))
    assert st.get_latency_samples() == 512 and st.get_latency_seconds(1000.0) == 0.512
    rt = sg.RealTimeStft(sg.StreamingStftConfig(frame_length=256, hop_length=128, center=False), 128, 2)
    assert rt.process_block(np.full(128, 0.5)) == 0
    for _ in range(4):
        assert rt.process_block(np.full(128, 0.5)) == 1
    assert rt.available_spectra_count() == 2 and rt.is_buffer_full()
    assert rt.get_statistics().base_statistics.frames_generated == 4
    with pytest.raises(ValueError_):
        rt.process_block(np.ones(5))
    assert len(rt.get_all_spectra()) == 2 and rt.get_spectrum() is None
    rt.reset()
    assert rt.get_statistics().base_statistics.samples_processed == 0


def test_phase_coupling_peaks(host_signal):
    """higher_order.rs:868-912 — discrete output (peak list), so it is compared on the CPU only: quadratic phase coupling
    between 12 and 20 cycles per 128 samples shows up as a bicoherence peak at that bin pair."""
    sg = host_signal
    t = np.arange(1024)
    x = (np.cos(2 * np.pi * 12 * t / 128 + 0.4) + np.cos(2 * np.pi * 20 * t / 128 + 1.1)
         + np.cos(2 * np.pi * 32 * t / 128 + 1.5) + 0.05 * np.random.default_rng(3).standard_normal(1024))
    got = sg.detect_phase_coupling(x, 128, "hann", 1.0, 0.3)
    ref = so.detect_phase_coupling(x, 128, "hann", 1.0, 0.3)
    assert len(got) > 0 and abs(len(got) - len(ref)) <= 4  # weak noise peaks can tie differently at rounding level
    # the strongest peaks are unambiguous: same places (a symmetric pair may swap order), same values
    assert {(p[0], p[1]) for p in got[:4]} == {(p[0], p[1]) for p in ref[:4]}
    for a, b in zip(got[:4], ref[:4]):
        assert abs(a[2] - b[2]) <= 1e-9 * abs(b[2])
    top = sorted((round(got[0][0] * 128), round(got[0][1] * 128)))
    assert abs(top[0] - 12) <= 1 and abs(top[1] - 20) <= 1


def test_spectral_host_logic_random_parameters(host_signal):
    """Seeded sweep over odd / tiny / mismatched parameters of the spectral callers (argument handling, segment
    counts, padding, axis construction) against the literal loops."""
    sg = host_signal
    rng = np.random.default_rng(77)
    wins = ["hann", "hamming", "blackman", "boxcar"]
    dets = ["constant", "linear", "none"]
    for _ in range(120):
        n = int(rng.integers(8, 400))
        x = rng.standard_normal(n) + 0.01 * np.arange(n)
        nperseg = int(rng.integers(2, min(n, 96) + 1))
        noverlap = int(rng.integers(0, nperseg))
        nfft = nperseg + int(rng.integers(0, 40)) if rng.random() < 0.5 else None
        fs = float(rng.choice([1.0, 7.5, 100.0]))
        w, d = str(rng.choice(wins)), str(rng.choice(dets))
        sc_ = str(rng.choice(["density", "spectrum"]))
        if nperseg < 3 and w != "boxcar":
            w = "boxcar"  # the reference's symmetric windows divide by nperseg - 1
        sc.compare(sg.welch(x, fs, w, nperseg, noverlap, nfft, d, sc_), so.welch(x, fs, w, nperseg, noverlap, nfft, d, sc_), 1e-10)
        b = str(rng.choice(["zeros", "extend", "none"]))
        pad = bool(rng.integers(0, 2))
        sc.compare(sg.stft(x, fs, w, nperseg, noverlap, nfft, d, b, pad), so.stft(x, fs, w, nperseg, noverlap, nfft, d, b, pad), 1e-10)
        pn = n + int(rng.integers(0, 30)) if rng.random() < 0.5 else None
        sc.compare(sg.periodogram(x, fs, w if n > 2 else "boxcar", pn, d, sc_), so.periodogram(x, fs, w if n > 2 else "boxcar", pn, d, sc_), 1e-10)
