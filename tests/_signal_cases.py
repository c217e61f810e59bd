"""Shared case list for the scirs2-signal caller tests: each entry runs the same call on the product
module (scirs_b200.signal) and on the oracle (oracle/signal_oracle.py) and returns both results."""
import numpy as np


def _sig(n, seed=0):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 100.0
    return np.sin(2 * np.pi * 10.0 * t) + 0.5 * np.sin(2 * np.pi * 23.0 * t + 0.3) + 0.1 * rng.standard_normal(n) + 0.2 * t


def cases():
    out = []

    def add(name, fn):
        out.append((name, fn))

    for n, kw in ((1000, {}), (1024, dict(window="hann")), (777, dict(window="blackman", nfft=900, detrend="linear",
                                                                       scaling="spectrum")),
                  (64, dict(window="hamming", detrend="none", fs=8.0))):
        x = _sig(n, n)
        add(f"periodogram-{n}", lambda sg, so, x=x, kw=kw: (sg.periodogram(x, kw.get("fs", 100.0), kw.get("window"), kw.get("nfft"), kw.get("detrend"), kw.get("scaling")),
                                                            so.periodogram(x, kw.get("fs", 100.0), kw.get("window"), kw.get("nfft"), kw.get("detrend"), kw.get("scaling"))))
    for n, args in ((2000, (100.0, None, 256, 128, None, None, None)),
                    (3001, (50.0, "hamming", 200, 37, 300, "linear", "spectrum")),
                    (500, (1.0, "boxcar", None, None, None, "none", None)),
                    (4096, (1.0, "blackman", 512, 448, 512, None, None))):
        x = _sig(n, n + 1)
        add(f"welch-{n}", lambda sg, so, x=x, a=args: (sg.welch(x, *a), so.welch(x, *a)))
    for n, args in ((2000, (1000.0, None, 128, 64, None, None, None, None)),
                    (1500, (10.0, "hamming", 100, 30, 160, "linear", "extend", True)),
                    (900, (1.0, "blackman", 64, 63, None, "none", "none", False)),
                    (700, (2.0, "boxcar", 50, 0, 64, "constant", "zeros", True))):
        x = _sig(n, n + 2)
        add(f"stft-{n}", lambda sg, so, x=x, a=args: (sg.stft(x, *a), so.stft(x, *a)))
    x = _sig(1000, 5)
    for mode in ("psd", "magnitude", "phase"):
        add(f"spectrogram-{mode}", lambda sg, so, x=x, m=mode: (sg.spectrogram(x, 100.0, None, 128, None, None, None, "spectrum" if m == "psd" else None, m),
                                                                so.spectrogram(x, 100.0, None, 128, None, None, None, "spectrum" if m == "psd" else None, m)))
    for n in (1000, 1024, 333):
        x = _sig(n, n + 3)
        add(f"wiener-{n}", lambda sg, so, x=x: (sg.wiener_filter(x), so.wiener_filter_freq(x)))
        add(f"wiener-snr-{n}", lambda sg, so, x=x: (sg.wiener_filter_freq(x, sg.WienerConfig(noise_power=0.02, prior_snr=3.0)),
                                                    so.wiener_filter_freq(x, 0.02, 3.0)))
        add(f"specsub-{n}", lambda sg, so, x=x: (sg.spectral_subtraction(x), so.spectral_subtraction(x)))
        add(f"specsub-given-{n}", lambda sg, so, x=x: (sg.spectral_subtraction(x, np.full(10, 0.3), 2.0, 0.05),
                                                       so.spectral_subtraction(x, np.full(10, 0.3), 2.0, 0.05)))
        add(f"psdwiener-{n}", lambda sg, so, x=x: (sg.psd_wiener_filter(x), so.psd_wiener_filter(x)))
        add(f"psdwiener-given-{n}", lambda sg, so, x=x: (sg.psd_wiener_filter(x, np.linspace(1, 2, 40), np.full(50, 0.5)),
                                                         so.psd_wiener_filter(x, np.linspace(1, 2, 40), np.full(50, 0.5))))

    def streaming(sg, so, L, hop, center, kw, n, block):
        x = _sig(n, L)
        a = sg.StreamingStft(sg.StreamingStftConfig(frame_length=L, hop_length=hop, center=center, **kw))
        b = so.StreamingStft(L, hop, kw.get("window", "hann"), center, kw.get("magnitude_only", False),
                             kw.get("log_magnitude", False), kw.get("power", 1.0))
        ra = a.process_batch(x, block) + a.flush()
        rb = b.process_batch(list(x), block) + b.flush()
        assert a.frames_generated == b.frames_generated
        return (np.array(ra),), (np.array(rb),)

    for L, hop, center, kw, n, block in ((128, 64, False, {}, 512, 64), (256, 128, True, dict(window="hamming"), 2000, 100),
                                         (100, 30, True, dict(magnitude_only=True, power=2.0, log_magnitude=True), 1000, 77),
                                         (96, 96, False, dict(magnitude_only=True, power=1.5, window="bartlett"), 700, 200)):
        add(f"streaming-{L}-{hop}", lambda sg, so, a=(L, hop, center, kw, n, block): streaming(sg, so, *a))

    for n, nfft, win in ((256, 64, "hann"), (1000, 128, None), (300, 100, "hamming")):
        x = _sig(n, n + 9) ** 2  # quadratic coupling, so the bispectrum is not numerically zero
        add(f"bispec-direct-{n}", lambda sg, so, x=x, nfft=nfft, win=win: (
            (sg.compute_bispectrum(x, sg.HigherOrderConfig(estimator="direct", nfft=nfft, window=win))[0],),
            (so.direct_bispectrum(x, nfft, win),)))
        add(f"bispec-welch-{n}", lambda sg, so, x=x, nfft=nfft, win=win: (
            (sg.compute_bispectrum(x, sg.HigherOrderConfig(estimator="welch", nfft=nfft, window=win))[0],),
            (so.welch_bispectrum(x, nfft, win),)))
        add(f"powerspec-{n}", lambda sg, so, x=x, nfft=nfft, win=win: (
            (sg.compute_power_spectrum(x, sg.HigherOrderConfig(nfft=nfft, window=win))[0],),
            (so.power_spectrum(x, nfft, win),)))

    def cqt_case(sg, so, cfgkw, n, seed, inverse):
        fs = cfgkw["fs"]
        t = np.arange(n) / fs
        x = np.sin(2 * np.pi * 440.0 * t) + 0.3 * np.sin(2 * np.pi * 700.0 * t) + 0.05 * np.random.default_rng(seed).standard_normal(n)
        cfg = sg.CqtConfig(**cfgkw)
        r = sg.constant_q_transform(x, cfg)
        q = cfg.q_factor if cfg.q_factor is not None else 1.0 / (2.0 ** (1.0 / cfg.bins_per_octave) - 1.0)
        kern, freqs, n_fft = so.cqt_kernel(cfg.f_min, cfg.f_max, cfg.bins_per_octave, q, fs, cfg.window_type,
                                           cfg.window_scaling, cfg.use_sparse)
        assert r.kernel.n_fft == n_fft and [len(k.indices) for k in r.kernel.kernels] == [len(i) for i, _ in kern]
        if cfg.hop_size is None:
            ref, times = so.cqt_frame(x, kern, n_fft).reshape(-1, 1), None
            got = (r.cqt, r.frequencies)
            exp = (ref, freqs)
        else:
            ref, times = so.cqt_spectrogram(x, kern, n_fft, fs, cfg.hop_size)
            got = (r.cqt, r.frequencies, r.times)
            exp = (ref, freqs, times)
        got += (sg.chromagram(r), sg.cqt_magnitude(r, True), sg.cqt_phase(r))
        exp += (so.chromagram(ref, freqs), 20.0 * np.log10(np.abs(ref) / (np.abs(ref).max() + 1e-10)), np.angle(ref))
        if inverse:
            got += (sg.inverse_constant_q_transform(r),)
            exp += (so.icqt(ref, kern, n_fft, fs, times),)
        return got, exp

    for name, kw, n, inverse in (
            ("frame-short", dict(f_min=220.0, f_max=880.0, bins_per_octave=12, fs=8000.0), 500, True),
            ("frame-chunks", dict(f_min=220.0, f_max=1760.0, bins_per_octave=12, fs=8000.0, window_type="hamming"), 3000, False),
            ("spec", dict(f_min=200.0, f_max=1600.0, bins_per_octave=6, fs=8000.0, hop_size=128, window_type="blackman"), 2000, True),
            ("dense", dict(f_min=300.0, f_max=1200.0, bins_per_octave=4, fs=8000.0, hop_size=100, use_sparse=False, q_factor=5.0,
                           window_scaling=1.5, window_type="bartlett"), 900, True)):
        add(f"cqt-{name}", lambda sg, so, a=(kw, n, len(name), inverse): cqt_case(sg, so, *a))

    for n in (1000, 1024, 255, 2, 1):
        x = _sig(max(n, 4), n + 20)[:n]
        add(f"hilbert-{n}", lambda sg, so, x=x: ((sg.hilbert(x), sg.envelope(x)), (so.hilbert(x), np.abs(so.hilbert(x)))))
    for n in (1000, 512):
        x = _sig(n, n + 21)
        add(f"instphase-{n}", lambda sg, so, x=x: ((sg.instantaneous_phase(x), sg.instantaneous_phase(x, True), sg.instantaneous_frequency(x, 100.0)),
                                                   (so.instantaneous_phase(x), so.instantaneous_phase(x, True), so.instantaneous_frequency(x, 100.0))))
    hann = lambda m: 0.5 * (1.0 - np.cos(2.0 * np.pi * np.arange(m) / (m - 1)))
    for name, n, kw in (("plain", 64, dict()), ("nopad-raw", 50, dict(analytic=False, zero_padding=False)),
                        ("sp-odd", 48, dict(zero_padding=False, time_window=hann(11), freq_window=hann(31))),
                        ("sp-even", 40, dict(time_window=hann(12), freq_window=hann(100))),
                        ("long-window", 33, dict(time_window=hann(81), freq_window=hann(200), zero_padding=False))):
        x = _sig(n, n + 22)
        add(f"wvd-{name}", lambda sg, so, x=x, kw=kw: (
            (sg.wigner_ville(x, sg.WvdConfig(**kw)),),
            (so.wigner_ville(x, kw.get("analytic", True), kw.get("zero_padding", True), kw.get("time_window"), kw.get("freq_window")),)))
    x, y = _sig(40, 1), _sig(40, 2)
    add("wvd-cross", lambda sg, so, x=x, y=y: ((sg.cross_wigner_ville(x, y, sg.WvdConfig(zero_padding=False)),),
                                     (so.cross_wvd(so.hilbert(x), so.hilbert(y), False),)))
    add("wvd-spwv", lambda sg, so, x=x: ((sg.smoothed_pseudo_wigner_ville(x, hann(9), hann(21), sg.WvdConfig(zero_padding=False)),),
                                    (so.wigner_ville(x, True, False, hann(9), hann(21)),)))
    for n, nfft, win in ((90, 32, "hann"), (200, 64, None), (61, 100, "hamming")):
        x = _sig(n, n + 23) ** 2
        add(f"bispec-indirect-{n}", lambda sg, so, x=x, nfft=nfft, win=win: (
            (sg.compute_bispectrum(x, sg.HigherOrderConfig(estimator="indirect", nfft=nfft, window=win))[0],
             sg.compute_triple_correlation(x, 9)),
            (so.indirect_bispectrum(x, nfft, win), so.triple_correlation(x, 9))))
    for n, nfft, win, fs in ((300, 64, None, 1.0), (500, 50, "hann", 8.0), (256, 33, "hamming", 2.0)):
        x = _sig(n, n + 24) ** 2
        add(f"bicoherence-{n}", lambda sg, so, x=x, a=(nfft, win, None, fs): ((sg.bicoherence(x, *a)[0],), (so.bicoherence(x, *a),)))

    for n, nfft, win in ((300, 64, "hann"), (200, 50, None)):
        x = _sig(n, n + 25) ** 2
        add(f"higher-order-rest-{n}", lambda sg, so, x=x, nfft=nfft, win=win: (
            (sg.trispectrum(x, nfft, win), sg.biamplitude(x, nfft, win)[0], sg.cumulative_bispectrum(x, nfft, win)[0],
             sg.cumulative_bispectrum(x, nfft, win)[1], sg.skewness_spectrum(x, nfft, win)[0]),
            (so.trispectrum(x, nfft, win), so.biamplitude(x, nfft, win), so.cumulative_bispectrum(x, nfft, win)[0],
             so.cumulative_bispectrum(x, nfft, win)[1], so.skewness_spectrum(x, nfft, win))))
    return out


def compare(got, ref, tol):
    got = got if isinstance(got, tuple) else (got,)
    ref = ref if isinstance(ref, tuple) else (ref,)
    assert len(got) == len(ref)
    for g, r in zip(got, ref):
        g, r = np.asarray(g), np.asarray(r)
        assert g.shape == r.shape, (g.shape, r.shape)
        d, s = np.linalg.norm((g - r).ravel()), np.linalg.norm(r.ravel())
        assert (d / s if s > 0 else d) <= tol, (d, s)
