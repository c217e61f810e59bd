"""TEST INFRASTRUCTURE ONLY — CPU restatement of the scirs2-signal callers of the FFT hot path
(SURVEY 8f rank 4).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

Written the way the reference is: one segment at a time, each through the oracle's own `fft` / `ifft`
(scirs2_fft_oracle.py, which restates fft/algorithms.rs including the next-power-of-two padding of
`fft(x, None)`), plain Python loops for the bin-by-bin passes.  Cited file:line ranges:

  get_window / apply_detrend / periodogram / welch / stft / spectrogram
                             scirs2-signal/src/spectral.rs:29-66, 77-117, 130-244, 257-410, 413-447, 468-628, 644-735
  estimate_noise_power / wiener_filter_freq / spectral_subtraction / smooth_psd / psd_wiener_filter
                             scirs2-signal/src/wiener.rs:709-739, 137-196, 439-546, 775-794, 560-655
  window (hann .. cosine)    scirs2-signal/src/window/mod.rs:38-107, 133-260, 580-590
  StreamingStft              scirs2-signal/src/streaming_stft.rs:124-438
  bispectrum (direct, Welch) scirs2-signal/src/higher_order.rs:250-332, 359-461

Pinning: the reference's unit tests for these modules hold no golden vectors (they assert peak positions,
shapes and value ranges: spectral.rs:743-935, streaming_stft.rs tests, wiener.rs tests); those assertions
are restated in tests/test_signal_oracle.py and cross-checked against scipy.signal where the two agree by
definition.  Beyond that: parity unpinned by the reference.
"""
from __future__ import annotations

import cmath
import math
from collections import deque
from typing import List, Optional

import numpy as np

from . import scirs2_fft_oracle as base

OracleError = base.OracleError


# ---- spectral.rs ---------------------------------------------------------------------------------

def get_window(window_type: str, nperseg: int) -> List[float]:  # spectral.rs:29-66
    w = window_type.lower()
    if w == "hann":
        return [0.5 * (1.0 - math.cos(2.0 * math.pi * i / (nperseg - 1))) for i in range(nperseg)]
    if w == "hamming":
        return [0.54 - 0.46 * math.cos(2.0 * math.pi * i / (nperseg - 1)) for i in range(nperseg)]
    if w == "blackman":
        return [0.42 - 0.5 * math.cos(2.0 * math.pi * i / (nperseg - 1))
                + 0.08 * math.cos(4.0 * math.pi * i / (nperseg - 1)) for i in range(nperseg)]
    if w in ("boxcar", "rectangular"):
        return [1.0] * nperseg
    raise OracleError(f"Unknown window type: {window_type}")


def apply_detrend(x: List[float], detrend_type: str) -> List[float]:  # spectral.rs:77-117
    if detrend_type == "constant":
        mean = sum(x) / len(x)
        return [v - mean for v in x]
    if detrend_type == "linear":
        n = len(x)
        sum_x = float(sum(range(n)))
        sum_y = sum(x)
        sum_xx = float(sum(i * i for i in range(n)))
        sum_xy = sum(i * y for i, y in enumerate(x))
        slope = (n * sum_xy - sum_x * sum_y) / (n * sum_xx - sum_x * sum_x)
        intercept = (sum_y - slope * sum_x) / n
        return [y - (slope * i + intercept) for i, y in enumerate(x)]
    if detrend_type == "none":
        return list(x)
    raise OracleError(f"Unknown detrend option: {detrend_type}")


def _fftfreq(n: int, d: float) -> List[float]:  # scirs2-fft/src/helper.rs fftfreq: k / (n d), negative half wrapped
    val = 1.0 / (n * d)
    half = (n - 1) // 2 + 1
    return [k * val for k in range(half)] + [-(n // 2 - k) * val for k in range(n // 2)]


def _segment_fft(seg: List[float], win: List[float], detrend: str, nfft: int) -> np.ndarray:
    d = apply_detrend(seg, detrend)
    padded = [a * w for a, w in zip(d, win)]
    if nfft > len(padded):
        padded = padded + [0.0] * (nfft - len(padded))
    return base.fft(np.array(padded, dtype=np.float64), None)


def periodogram(x, fs=None, window=None, nfft=None, detrend=None, scaling=None):  # spectral.rs:130-244
    x = [float(v) for v in x]
    if not x:
        raise OracleError("Input array is empty")
    fs = 1.0 if fs is None else fs
    nfft = len(x) if nfft is None else nfft
    window = window or "boxcar"
    detrend = detrend or "constant"
    scaling = scaling or "density"
    if fs <= 0.0:
        raise OracleError("Sampling frequency must be positive")
    if nfft < len(x):
        raise OracleError("NFFT must be at least as large as signal length")
    win = get_window(window, len(x))
    scale = 1.0 / sum(w * w for w in win)
    spectrum = _segment_fft(x, win, detrend, nfft)
    pg = [(c.real * c.real + c.imag * c.imag) * scale / (fs * len(x)) for c in spectrum]
    freqs = _fftfreq(nfft, 1.0 / fs)
    n_half = nfft // 2 + nfft % 2
    rf, rp = [], []
    for i in range(n_half):
        if i < len(freqs) and i < len(pg):
            rf.append(freqs[i])
            rp.append(pg[i] if scaling == "density" else pg[i] * fs)
    return np.array(rf), np.array(rp)


def welch(x, fs=None, window=None, nperseg=None, noverlap=None, nfft=None, detrend=None, scaling=None):
    # spectral.rs:257-410
    x = [float(v) for v in x]
    if not x:
        raise OracleError("Input array is empty")
    fs = 1.0 if fs is None else fs
    nperseg = min(256, len(x)) if nperseg is None else nperseg
    noverlap = nperseg // 2 if noverlap is None else noverlap
    nfft = nperseg if nfft is None else nfft
    window = window or "hann"
    detrend = detrend or "constant"
    scaling = scaling or "density"
    if fs <= 0.0 or nfft < nperseg or noverlap >= nperseg:
        raise OracleError("bad parameters")
    win = get_window(window, nperseg)
    scale = 1.0 / sum(w * w for w in win)
    step = nperseg - noverlap
    num_segments = (len(x) - noverlap) // step if len(x) >= noverlap else 0
    if num_segments < 1:
        raise OracleError("Not enough data points for given nperseg and noverlap")
    n_half = nfft // 2 + nfft % 2
    freqs = _fftfreq(nfft, 1.0 / fs)[:n_half]
    avg = [0.0] * n_half
    for i in range(num_segments):
        start, end = i * step, i * step + nperseg
        if end > len(x):
            break
        spectrum = _segment_fft(x[start:end], win, detrend, nfft)
        for j, c in enumerate(spectrum[:n_half]):
            avg[j] += (c.real * c.real + c.imag * c.imag) * scale / (fs * nperseg)
    avg = [p / num_segments for p in avg]
    if scaling != "density":
        avg = [p * fs for p in avg]
    return np.array(freqs), np.array(avg)


def apply_boundary(x: List[float], nperseg: int, boundary: str) -> List[float]:  # spectral.rs:413-447
    pad = nperseg // 2
    if boundary == "zeros":
        return [0.0] * pad + list(x) + [0.0] * pad
    if boundary == "extend":
        return [x[0]] * pad + list(x) + [x[-1]] * pad
    if boundary == "none":
        return list(x)
    raise OracleError(f"Unknown boundary option: {boundary}")


def stft(x, fs=None, window=None, nperseg=None, noverlap=None, nfft=None, detrend=None, boundary=None, padded=None):
    # spectral.rs:468-628; returns (freqs, times, Z[segment][bin])
    x = [float(v) for v in x]
    if not x:
        raise OracleError("Input array is empty")
    fs = 1.0 if fs is None else fs
    nperseg = min(256, len(x)) if nperseg is None else nperseg
    noverlap = nperseg // 2 if noverlap is None else noverlap
    nfft = nperseg if nfft is None else nfft
    window = window or "hann"
    detrend = detrend or "constant"
    boundary = boundary or "zeros"
    padded = True if padded is None else padded
    if fs <= 0.0 or nfft < nperseg or noverlap >= nperseg:
        raise OracleError("bad parameters")
    win = get_window(window, nperseg)
    sig = apply_boundary(x, nperseg, boundary) if padded else x
    step = nperseg - noverlap
    num_segments = (len(sig) - noverlap) // step if len(sig) >= noverlap else 0
    if num_segments < 1:
        raise OracleError("Not enough data points for given nperseg and noverlap")
    n_half = nfft // 2 + nfft % 2
    freqs = _fftfreq(nfft, 1.0 / fs)[:n_half]
    times = [(i * step + nperseg // 2) / fs for i in range(num_segments)]
    out = [[0j] * num_segments for _ in range(n_half)]
    for i in range(num_segments):
        start, end = i * step, i * step + nperseg
        if end > len(sig):
            break
        spectrum = _segment_fft(sig[start:end], win, detrend, nfft)
        for j, v in enumerate(spectrum[:n_half]):
            out[j][i] = complex(v)
    Z = [[out[j][i] for j in range(n_half)] for i in range(num_segments)]
    return np.array(freqs), np.array(times), np.array(Z, dtype=np.complex128).reshape(num_segments, n_half)


def spectrogram(x, fs=None, window=None, nperseg=None, noverlap=None, nfft=None, detrend=None, scaling=None,
                mode=None):  # spectral.rs:644-735
    mode = mode or "psd"
    scaling = scaling or "density"
    freqs, times, Z = stft(x, fs, window, nperseg, noverlap, nfft, detrend, "zeros", True)
    if mode == "psd":
        fsv = 1.0 if fs is None else fs
        npg = min(256, len(x)) if nperseg is None else nperseg
        win = get_window(window or "hann", npg)
        scale = 1.0 / sum(w * w for w in win)
        S = np.array([[(c.real ** 2 + c.imag ** 2) * scale / (fsv * npg) * (1.0 if scaling == "density" else fsv)
                       for c in col] for col in Z])
    elif mode == "magnitude":
        S = np.array([[abs(c) for c in col] for col in Z])
    elif mode in ("angle", "phase"):
        S = np.array([[cmath.phase(c) for c in col] for col in Z])
    else:
        raise OracleError(f"mode {mode}")
    return freqs, times, S


# ---- wiener.rs -----------------------------------------------------------------------------------

def _median(v: List[float]) -> float:
    n = len(v)
    return (v[n // 2 - 1] + v[n // 2]) / 2.0 if n % 2 == 0 else v[n // 2]


def estimate_noise_power(signal) -> float:  # wiener.rs:709-739
    values = sorted(float(v) for v in signal)
    median = _median(values)
    dev = sorted(abs(v - median) for v in values)
    return (1.4826 * _median(dev)) ** 2


def wiener_filter_freq(signal, noise_power=None, prior_snr=None, regularization=1e-10) -> np.ndarray:
    # wiener.rs:137-196
    s = np.asarray(signal, dtype=np.float64)
    n = s.size
    noise = noise_power if noise_power is not None else estimate_noise_power(s)
    spec = base.fft(s, None)
    out = []
    for c in spec:
        power = c.real * c.real + c.imag * c.imag
        snr = 1.0 if prior_snr is None else prior_snr
        out.append(c * (power / (power + snr * noise + regularization)))
    return np.array([c.real for c in base.ifft(np.array(out), None)[:n]])


def spectral_subtraction(signal, noise_power=None, alpha=None, beta=None) -> np.ndarray:  # wiener.rs:439-546
    s = np.asarray(signal, dtype=np.float64)
    n = s.size
    a = 1.0 if alpha is None else alpha
    b = 0.01 if beta is None else beta
    spec = [complex(c) for c in base.fft(s, None)]
    if noise_power is not None:
        noise = [float(v) for v in noise_power]
    else:
        ns = int(min(n * 0.05, 100.0))
        if ns < 4:
            raise OracleError("Signal too short to estimate noise spectrum")
        nf = base.fft(s[:ns], n)
        noise = [(c.real ** 2 + c.imag ** 2) / n for c in nf[: n // 2 + 1]]
    for i in range(n // 2 + 1):
        mag, phase = abs(spec[i]), cmath.phase(spec[i])
        npw = noise[i] if i < len(noise) else noise[-1]
        new_mag = math.sqrt(max(mag ** 2 - a * npw, b * mag ** 2))
        spec[i] = cmath.rect(new_mag, phase)
        if 0 < i < n // 2:
            spec[n - i] = cmath.rect(new_mag, -phase)
    return np.array([c.real for c in base.ifft(np.array(spec), None)[:n]])


def smooth_psd(psd: List[float]) -> List[float]:  # wiener.rs:775-794
    n = len(psd)
    window_size = int(min(max(n * 0.02, 3.0), 15.0))
    half = window_size // 2
    out = []
    for i in range(n):
        st, en = max(i - half, 0), min(i + half + 1, n)
        out.append(sum(psd[st:en]) / (en - st))
    return out


def psd_wiener_filter(signal, signal_psd=None, noise_psd=None) -> np.ndarray:  # wiener.rs:560-655
    s = np.asarray(signal, dtype=np.float64)
    n = s.size
    spec = [complex(c) for c in base.fft(s, None)]
    if signal_psd is not None:
        s_psd = [float(v) for v in signal_psd]
    else:
        s_psd = smooth_psd([(spec[i].real ** 2 + spec[i].imag ** 2) / n for i in range(n // 2 + 1)])
    n_psd = [float(v) for v in noise_psd] if noise_psd is not None else [estimate_noise_power(s)] * (n // 2 + 1)
    for i in range(n // 2 + 1):
        mag, phase = abs(spec[i]), cmath.phase(spec[i])
        sp = s_psd[i] if i < len(s_psd) else 0.0
        npw = n_psd[i] if i < len(n_psd) else 0.0
        gain = sp / (sp + npw) if sp + npw > 1e-10 else 0.0
        spec[i] = cmath.rect(mag * gain, phase)
        if 0 < i < n // 2:
            spec[n - i] = cmath.rect(mag * gain, -phase)
    return np.array([c.real for c in base.ifft(np.array(spec), None)[:n]])


# ---- window/mod.rs (subset) and streaming_stft.rs ------------------------------------------------

def signal_window(window_type: str, length: int, periodic: bool) -> List[float]:  # window/mod.rs:38-107 ...
    if length == 0:
        raise OracleError("Window length must be positive")
    if length <= 1:
        return [1.0] * length
    n = length if not periodic else length + 1  # _extend(m, sym = !periodic)
    w = window_type.lower()
    if w in ("hann", "hanning"):
        v = [0.5 * (1.0 - math.cos(2.0 * math.pi * i / (n - 1))) for i in range(n)]
    elif w == "hamming":
        v = [0.54 - 0.46 * math.cos(2.0 * math.pi * i / (n - 1)) for i in range(n)]
    elif w == "blackman":
        v = [0.42 - 0.5 * math.cos(2.0 * math.pi * i / (n - 1)) + 0.08 * math.cos(4.0 * math.pi * i / (n - 1))
             for i in range(n)]
    elif w == "bartlett":
        m2 = (n - 1) / 2.0
        v = [1.0 - abs((i - m2) / m2) for i in range(n)]
    elif w == "cosine":
        v = [math.sin(math.pi * i / (n - 1)) for i in range(n)]
    elif w in ("boxcar", "rectangular"):
        v = [1.0] * n
    else:
        raise OracleError(f"Unknown window type: {window_type}")
    return v[:length]  # _truncate


class StreamingStft:  # streaming_stft.rs:124-438
    def __init__(self, frame_length=512, hop_length=256, window="hann", center=True, magnitude_only=False,
                 log_magnitude=False, power=1.0, log_epsilon=1e-10):
        self.L, self.hop, self.center = frame_length, hop_length, center
        self.magnitude_only, self.log_magnitude, self.power, self.eps = magnitude_only, log_magnitude, power, log_epsilon
        self.window = signal_window(window, frame_length, True)
        self.buf = deque([0.0] * (frame_length // 2) if center else [])
        self.frames_generated = 0

    def _fft(self, frame):
        spec = base.fft(np.array(frame, dtype=np.float64), None)[: self.L // 2 + 1]
        if not self.magnitude_only:
            return np.array(spec)
        if self.power == 1.0:
            m = [abs(c) for c in spec]
        elif self.power == 2.0:
            m = [c.real ** 2 + c.imag ** 2 for c in spec]
        else:
            m = [abs(c) ** self.power for c in spec]
        if self.log_magnitude:
            m = [math.log(v + self.eps) for v in m]
        return np.array(m, dtype=np.complex128)

    def process_frame(self, samples):
        self.buf.extend(float(v) for v in samples)
        if len(self.buf) < self.L:
            return None
        frame = [self.buf[i] * self.window[i] for i in range(self.L)]
        out = self._fft(frame)
        for _ in range(self.hop):
            if self.buf:
                self.buf.popleft()
        self.frames_generated += 1
        return out

    def process_batch(self, data, frame_size):
        res, start = [], 0
        while start + frame_size <= len(data):
            r = self.process_frame(data[start:start + frame_size])
            if r is not None:
                res.append(r)
            start += frame_size
        if start < len(data):
            r = self.process_frame(data[start:])
            if r is not None:
                res.append(r)
        return res

    def flush(self):
        res = []
        while len(self.buf) >= self.hop:
            frame = [0.0] * self.L
            for i in range(min(len(self.buf), self.L)):
                frame[i] = self.buf[i]
            res.append(self._fft([f * w for f, w in zip(frame, self.window)]))
            for _ in range(self.hop):
                if self.buf:
                    self.buf.popleft()
            self.frames_generated += 1
        return res


# ---- higher_order.rs -----------------------------------------------------------------------------

def direct_bispectrum(signal, nfft: int, window: Optional[str]) -> np.ndarray:  # higher_order.rs:289-332
    s = np.asarray(signal, dtype=np.float64)
    if window is not None:
        s = s * np.array(signal_window(window, s.size, True))
    X = base.fft(s, nfft)
    nb = nfft // 2 + 1
    B = np.zeros((nb, nb), dtype=np.complex128)
    for i in range(nb):
        for j in range(i + 1):
            v = X[i] * X[j] * np.conj(X[(i + j) % nfft])
            B[i, j] = v
            B[j, i] = v
    return B


def welch_bispectrum(signal, nfft: int, window: Optional[str], overlap: float = 0.5,
                     n_segments: Optional[int] = None) -> np.ndarray:  # higher_order.rs:359-431
    s = np.asarray(signal, dtype=np.float64)
    n = s.size
    size = min(nfft, n)
    ov = int(round(size * overlap))  # f64::round is half away from zero; equal for the .5 cases used in tests
    step = size - ov
    if step == 0:
        raise OracleError("Overlap too large")
    nseg = n_segments if n_segments is not None else int(math.floor((n - ov) / step))
    if nseg == 0:
        raise OracleError("Signal too short")
    nb = nfft // 2 + 1
    acc = np.zeros((nb, nb), dtype=np.complex128)
    for i in range(nseg):
        st = i * step
        en = min(st + size, n)
        if en - st < 4:
            continue
        acc += direct_bispectrum(s[st:en], nfft, window)
    return acc / nseg


def power_spectrum(signal, nfft: int, window: Optional[str]) -> np.ndarray:  # higher_order.rs:464-508
    s = np.asarray(signal, dtype=np.float64)
    if window is not None:
        s = s * np.array(signal_window(window, s.size, True))
    X = base.fft(s, nfft)
    nb = nfft // 2 + 1
    p = np.array([(X[i].real ** 2 + X[i].imag ** 2) / nfft for i in range(nb)])
    if nb > 2:
        p[1:nb - 1] *= 2.0
    return p


# ---- cqt.rs --------------------------------------------------------------------------------------

def _odd_ceil(v: float) -> int:
    k = int(math.ceil(v))
    return k + 1 if k % 2 == 0 else k


def _pow2(n: int) -> int:  # cqt.rs:496-502
    p = 1
    while p < n:
        p *= 2
    return p


def cqt_kernel(f_min, f_max, bins_per_octave, q, fs, window_type="hann", window_scaling=None, use_sparse=True):
    """cqt.rs:229-347 -> (list of (indices, values), frequencies, n_fft)."""
    n_bins = int(math.ceil(math.log2(f_max / f_min) * bins_per_octave))
    freqs = [f_min * 2.0 ** (k / bins_per_octave) for k in range(n_bins)]
    ws = 1.0 if window_scaling is None else window_scaling
    n_fft = _pow2(_odd_ceil(ws * q * fs / f_min))
    kernels = []
    for f in freqs:
        klen = _odd_ceil(ws * q * fs / f)
        if window_type.lower() not in ("hann", "hanning", "hamming", "blackman", "bartlett", "rectangular", "boxcar"):
            raise OracleError(f"Unsupported window type: {window_type}")
        win = signal_window(window_type, klen, False)
        center = (klen - 1) / 2.0
        vals = []
        for n in range(klen):
            ph = 2.0 * math.pi * f * ((n - center) / fs)
            vals.append(complex(math.cos(ph), math.sin(ph)) * win[n])
        norm = math.sqrt(sum(v.real ** 2 + v.imag ** 2 for v in vals))
        padded = np.zeros(n_fft, dtype=np.complex128)
        for n in range(klen):
            padded[n] = vals[n] / norm
        K = base.fft(padded, None)
        if use_sparse:
            idx = [i for i, v in enumerate(K) if abs(v) > 1e-6]
        else:
            idx = list(range(n_fft))
        kernels.append((np.array(idx, dtype=np.int64), np.array([K[i] for i in idx], dtype=np.complex128)))
    return kernels, np.array(freqs), n_fft


def cqt_frame(signal, kernels, n_fft) -> np.ndarray:  # cqt.rs:351-428
    s = np.asarray(signal, dtype=np.float64)
    n_chunks = 1 if s.size < n_fft else int(math.ceil(s.size / n_fft))
    acc = np.zeros(len(kernels), dtype=np.complex128)
    for c in range(n_chunks):
        padded = np.zeros(n_fft, dtype=np.complex128)
        seg = s[c * n_fft: min((c + 1) * n_fft, s.size)]
        padded[: seg.size] = seg
        X = base.fft(padded, None)
        for k, (idx, vals) in enumerate(kernels):
            acc[k] += np.sum(X[idx] * np.conj(vals)) / 1.0
    return acc / n_chunks


def cqt_spectrogram(signal, kernels, n_fft, fs, hop):  # cqt.rs:430-478 -> (cqt[n_bins][n_frames], times)
    s = np.asarray(signal, dtype=np.float64)
    n_frames = int(math.ceil(s.size / hop))
    out = np.zeros((len(kernels), n_frames), dtype=np.complex128)
    times = np.zeros(n_frames)
    for f in range(n_frames):
        st = f * hop
        en = min(st + n_fft, s.size)
        times[f] = (st + (en - st) // 2) / fs
        frame = np.zeros(n_fft)
        frame[: en - st] = s[st:en]
        out[:, f] = cqt_frame(frame, kernels, n_fft)
    return out, times


def icqt(cq, kernels, n_fft, fs, times=None, target_length=None) -> np.ndarray:  # cqt.rs:609-700
    n_bins, n_frames = cq.shape
    if n_frames > 1 and times is not None:
        hop = int(math.floor((times[1] - times[0]) * fs + 0.5)) if len(times) > 1 else n_fft // 2
    else:
        hop = n_fft
    out_len = target_length if target_length is not None else ((n_frames - 1) * hop + n_fft if n_frames > 1 else n_fft)
    out = np.zeros(out_len)
    for f in range(n_frames):
        spec = np.zeros(n_fft, dtype=np.complex128)
        for b in range(n_bins):
            idx, vals = kernels[b]
            spec[idx] += cq[b, f] * vals
        sig = base.ifft(spec, None)
        st = f * hop
        for i in range(st, min(st + n_fft, out_len)):
            out[i] += sig[i - st].real
    peak = max((abs(v) for v in out), default=0.0)
    return out / peak if peak > 0.0 else out


def chromagram(cq, freqs, n_chroma=12, ref_note=0) -> np.ndarray:  # cqt.rs:713-760
    ref = ref_note % n_chroma
    chroma = np.zeros((n_chroma, cq.shape[1]))
    for i, f in enumerate(freqs):
        midi = int(69.0 + 12.0 * math.log2(f / 440.0))  # `as isize` truncates toward zero
        r = int(math.fmod(midi, n_chroma))
        b = int(math.fmod(r + n_chroma - ref, n_chroma))
        if 0 <= b < n_chroma:
            chroma[b] += np.abs(cq[i])
    for j in range(cq.shape[1]):
        t = chroma[:, j].sum()
        if t > 0.0:
            chroma[:, j] /= t
    return chroma


# ---- higher_order.rs, continued ------------------------------------------------------------------

def triple_correlation(signal, size: int) -> np.ndarray:  # higher_order.rs:511-556
    s = [float(v) for v in signal]
    n = len(s)
    mean = sum(s) / n
    c = [v - mean for v in s]
    max_lag = min(size, n // 3)
    if max_lag < 2:
        raise OracleError("Signal too short for triple correlation calculation")
    out = np.zeros((2 * max_lag - 1, 2 * max_lag - 1))
    for i in range(n):
        for t1 in range(max_lag):
            if i + t1 >= n:
                continue
            for t2 in range(max_lag):
                if i + t2 >= n:
                    continue
                out[t1 + max_lag - 1, t2 + max_lag - 1] += c[i] * c[i + t1] * c[i + t2]
    return out / n


def fft_2d(matrix: np.ndarray, nfft: int) -> np.ndarray:  # higher_order.rs:559-635
    rows, cols = matrix.shape
    nb = nfft // 2 + 1
    row_fft = []
    for i in range(rows):
        row = list(matrix[i].astype(np.complex128))
        if len(row) < nfft:
            row = row + [0j] * (nfft - len(row))
        row_fft.append(base.fft(np.array(row), None))
    res = np.zeros((nb, nb), dtype=np.complex128)
    for j in range(nb):
        col = [row_fft[i][j] for i in range(rows)]
        if len(col) < nfft:
            col = col + [0j] * (nfft - len(col))
        f = base.fft(np.array(col), None)
        for i in range(nb):
            res[i, j] = f[i]
    return res


def indirect_bispectrum(signal, nfft: int, window: Optional[str]) -> np.ndarray:  # higher_order.rs:334-356
    s = np.asarray(signal, dtype=np.float64)
    if window is not None:
        s = s * np.array(signal_window(window, s.size, True))
    return fft_2d(triple_correlation(s, nfft // 2 + 1), nfft)


def _round(v: float) -> int:  # f64::round, half away from zero
    return int(math.floor(v + 0.5)) if v >= 0 else -int(math.floor(-v + 0.5))


def bicoherence(signal, nfft: int, window: Optional[str] = None, n_segments=None, fs: float = 1.0) -> np.ndarray:
    # higher_order.rs:192-247
    B = welch_bispectrum(signal, nfft, window, 0.5, n_segments)
    P = power_spectrum(signal, nfft, window)
    nb = nfft // 2 + 1
    axis = np.linspace(0.0, fs / 2.0, nb)
    out = np.zeros((nb, nb))
    for i in range(nb):
        ii = _round(axis[i] * nfft / fs)
        for j in range(nb):
            jj = _round(axis[j] * nfft / fs)
            ss = _round((axis[i] + axis[j]) * nfft / fs) % nfft
            if ii < len(P) and jj < len(P) and ss < len(P):
                nf = math.sqrt(P[ii] * P[jj] * P[ss])
                if nf > 1e-10:
                    out[i, j] = abs(B[i, j]) / nf
    return out


# ---- hilbert.rs ----------------------------------------------------------------------------------

def hilbert(x) -> np.ndarray:  # scirs2-signal hilbert.rs:58-150
    s = np.asarray(x, dtype=np.float64)
    n = s.size
    if n == 0:
        raise OracleError("Input array is empty")
    spectrum = base.fft(s, None)
    h = [1 + 0j] * n
    if n % 2 == 0:
        h[0] = 1 + 0j
        h[n // 2] = 1 + 0j
        for i in range(1, n // 2):
            h[i] = -2j
        for i in range(n // 2 + 1, n):
            h[i] = 0j
    else:
        h[0] = 1 + 0j
        for i in range(1, (n + 1) // 2):
            h[i] = -2j
        for i in range((n + 1) // 2, n):
            h[i] = 0j
    filtered = np.array([sv * hv for sv, hv in zip(spectrum, h)])  # zip stops at n
    buf = base._process(filtered.astype(np.complex128).copy(), True)  # rustfft inverse, unscaled
    return buf * (1.0 / n)


def unwrap_phase(phase) -> np.ndarray:  # hilbert.rs:258-279
    out = [phase[0]]
    prev = phase[0]
    for p in phase[1:]:
        d = p - prev
        while d > math.pi:
            d -= 2.0 * math.pi
        while d < -math.pi:
            d += 2.0 * math.pi
        out.append(out[-1] + d)
        prev = p
    return np.array(out)


def instantaneous_phase(x, unwrap=False) -> np.ndarray:  # hilbert.rs:322-366
    a = hilbert(x)
    ph = np.array([math.atan2(c.imag, c.real) for c in a])
    return unwrap_phase(ph) if unwrap else ph


def instantaneous_frequency(x, fs) -> np.ndarray:  # hilbert.rs:228-292
    u = instantaneous_phase(x, True)
    f = [fs * (u[1] - u[0]) / (2.0 * math.pi)]
    for i in range(1, len(u) - 1):
        f.append(fs * (u[i + 1] - u[i - 1]) / (4.0 * math.pi))
    f.append(fs * (u[-1] - u[-2]) / (2.0 * math.pi))
    return np.array(f)


# ---- wvd.rs --------------------------------------------------------------------------------------

def cross_wvd(s1, s2, zero_padding=True, time_window=None, freq_window=None) -> np.ndarray:  # wvd.rs:232-344
    n = len(s1)
    n_fft = 2 * n if zero_padding else n
    out = np.zeros((n_fft // 2 + 1, n), dtype=np.complex128)
    tw = None
    if time_window is not None:
        w = list(time_window)
        if len(w) % 2 == 0:
            half = len(w) // 2
            tw = [0.0] * (len(w) + 1)
            for i in range(len(w)):
                tw[i + (1 if i >= half else 0)] = w[i]
        else:
            tw = w
    fw = None
    if freq_window is not None:
        w = list(freq_window)
        if len(w) < n_fft:
            fw = [0.0] * n_fft
            off = (n_fft - len(w)) // 2
            for i in range(len(w)):
                fw[i + off] = w[i]
        elif len(w) > n_fft:
            off = (len(w) - n_fft) // 2
            fw = w[off:off + n_fft]
        else:
            fw = w
    for t in range(n):
        acorr = np.zeros(n_fft, dtype=np.complex128)
        whl = len(tw) // 2 if tw is not None else n // 2
        for tau in range(-min(t, whl), min(n - t, whl + 1)):
            idx = tau + n_fft // 2
            if tw is not None:
                wi = tau + whl
                wv = tw[wi] if wi < len(tw) else 0.0
            else:
                wv = 1.0
            i1, i2 = t + tau, t - tau
            if 0 <= i1 < n and 0 <= i2 < n:
                acorr[idx] = s1[i1] * np.conj(s2[i2]) * wv
        if fw is not None:
            for i in range(n_fft):
                acorr[i] *= fw[i]
        spectrum = base.fft(acorr, None)
        for k in range(n_fft // 2 + 1):
            out[k, t] = spectrum[k]
    return out


def wigner_ville(signal, analytic=True, zero_padding=True, time_window=None, freq_window=None) -> np.ndarray:
    # wvd.rs:79-90, 192-229
    a = hilbert(signal) if analytic else np.asarray(signal, dtype=np.float64).astype(np.complex128)
    return cross_wvd(a, a, zero_padding, time_window, freq_window).real


# ---- higher_order.rs: the remaining estimators ----------------------------------------------------

def _wfft(signal, nfft, window):
    s = np.asarray(signal, dtype=np.float64)
    if window is not None:
        s = s * np.array(signal_window(window, s.size, True))
    return base.fft(s, nfft)


def trispectrum(signal, nfft, window=None) -> np.ndarray:  # higher_order.rs:638-684
    X = _wfft(signal, nfft, window)
    nb = nfft // 2 + 1
    out = np.zeros((nb, nb))
    for i in range(nb):
        for j in range(nb):
            out[i, j] = abs(X[i] * X[j] * np.conj(X[i]) * np.conj(X[j]))
    return out


def biamplitude(signal, nfft, window=None) -> np.ndarray:  # higher_order.rs:698-745
    X = _wfft(signal, nfft, window)
    nb = nfft // 2 + 1
    out = np.zeros((nb, nb))
    for i in range(nb):
        for j in range(nb):
            k = (i + j) % nfft
            if k < nb:
                out[i, j] = abs(X[i]) * abs(X[j]) * abs(X[k])
    return out


def cumulative_bispectrum(signal, nfft, window=None):  # higher_order.rs:762-804
    mag = np.abs(welch_bispectrum(signal, nfft, window))
    nb = mag.shape[0]
    bandwidth = np.linspace(1.0, float(nb // 2), 10)
    out = np.zeros(10)
    for i, bw in enumerate(bandwidth):
        b = _round(bw)
        if b > 0:
            tot, cnt = 0.0, 0
            for i1 in range(min(b, nb)):
                for i2 in range(min(b, nb)):
                    tot += mag[i1, i2]
                    cnt += 1
            if cnt > 0:
                out[i] = tot / cnt
    return out, bandwidth


def skewness_spectrum(signal, nfft, window=None) -> np.ndarray:  # higher_order.rs:818-852
    B = welch_bispectrum(signal, nfft, window)
    P = power_spectrum(signal, nfft, window)
    nb = nfft // 2 + 1
    out = np.zeros(nb)
    for i in range(nb):
        if P[i] > 1e-10:
            out[i] = abs(B[i, i]) / P[i] ** 1.5
    return out


def detect_phase_coupling(signal, nfft, window=None, fs=1.0, threshold=None):  # higher_order.rs:868-912
    thresh = 0.5 if threshold is None else threshold
    b = bicoherence(signal, nfft, window, None, fs)
    axis = np.linspace(0.0, fs / 2.0, nfft // 2 + 1)
    peaks = []
    for i in range(1, b.shape[0] - 1):
        for j in range(1, b.shape[1] - 1):
            v = b[i, j]
            if v > thresh:
                nbrs = [b[i - 1, j], b[i + 1, j], b[i, j - 1], b[i, j + 1], b[i - 1, j - 1], b[i + 1, j + 1],
                        b[i - 1, j + 1], b[i + 1, j - 1]]
                if all(v >= q for q in nbrs):
                    peaks.append((axis[i], axis[j], v))
    peaks.sort(key=lambda p: -p[2])
    return peaks
