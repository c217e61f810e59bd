"""Downstream callers of the FFT hot path in scirs2-signal, re-stated over this package.

The reference functions are host code that loops over segments and calls ``scirs2_fft::fft`` once per
segment.  Here the segment loop becomes ONE batched device transform: for spectral.rs the framing, detrending,
windowing, transform and (welch) the |X|^2 average all run on the device (``sfc_signal_spectra``); the other
callers build their frame matrix on the host and hand it to ``rfft_batch`` / ``fftn``.  Every FFT goes through the C ABI
(there is no CPU transform in this file; without the CUDA library every function raises).

Reference behaviour that is reproduced as written, because a drop-in must return the same numbers:

* ``scirs2_fft::fft(x, None)`` pads to the next power of two (fft/algorithms.rs:131-176), so
  ``periodogram`` / ``welch`` / ``stft`` of a non-power-of-two ``nfft`` return the first
  ``nfft/2 + nfft%2`` bins of a LONGER transform while the frequency axis is built for ``nfft``
  (spectral.rs:201-219, 376-380, 604-608).
* ``spectral_subtraction`` / ``psd_wiener_filter`` mirror bins around ``n`` (the signal length), not
  around the padded length (wiener.rs:497-529, 604-638).
* the spectral-density results are NOT doubled for the one-sided half (spectral.rs:203-207).

Covered: spectral.rs (periodogram, welch, stft, spectrogram), wiener.rs (the frequency-domain filters),
streaming_stft.rs (StreamingStft, RealTimeStft), higher_order.rs (direct and Welch bispectrum, power spectrum),
cqt.rs (constant-Q kernel, frame, spectrogram, inverse, chromagram), hilbert.rs (analytic signal, envelope,
instantaneous phase / frequency), wvd.rs (Wigner-Ville, cross and smoothed pseudo distributions).
"""
from __future__ import annotations

from collections import deque
from dataclasses import dataclass
from itertools import islice
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .error import ValueError_, check
from .fft import _ptr, fft, fftn, ifft, ifftn, rfft_batch


# ------------------------------------------------------------------------------------------------
# spectral.rs
# ------------------------------------------------------------------------------------------------

def spectral_window(window_type: str, nperseg: int) -> np.ndarray:
    """The private ``get_window`` of spectral.rs:29-66 (symmetric hann / hamming / blackman / boxcar)."""
    w = window_type.lower()
    i = np.arange(nperseg, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        if w == "hann":
            return 0.5 * (1.0 - np.cos(2.0 * np.pi * i / (nperseg - 1)))
        if w == "hamming":
            return 0.54 - 0.46 * np.cos(2.0 * np.pi * i / (nperseg - 1))
        if w == "blackman":
            return (0.42 - 0.5 * np.cos(2.0 * np.pi * i / (nperseg - 1))
                    + 0.08 * np.cos(4.0 * np.pi * i / (nperseg - 1)))
    if w in ("boxcar", "rectangular"):
        return np.ones(nperseg)
    raise ValueError_(f"Unknown window type: {window_type}")


def _fftfreq_head(nfft: int, fs: float, count: int) -> np.ndarray:
    """First ``count`` entries of ``helper::fftfreq(nfft, 1/fs)`` (helper.rs; count <= ceil(nfft/2), all non-negative)."""
    return np.arange(count, dtype=np.float64) / (nfft * (1.0 / fs))


def _next_pow2(n: int) -> int:
    p = 1
    while p < n:
        p <<= 1
    return p


def _check_common(fs: float, nfft: int, nperseg: int, noverlap: Optional[int]) -> None:
    if fs <= 0.0:
        raise ValueError_(f"Sampling frequency must be positive, got {fs}")
    if nfft < nperseg:
        raise ValueError_(f"nfft must be at least as large as nperseg, got {nfft} < {nperseg}")
    if noverlap is not None and noverlap >= nperseg:
        raise ValueError_(f"noverlap must be less than nperseg, got {noverlap} >= {nperseg}")


_DETREND = {"none": 0, "constant": 1, "linear": 2}


def _segment_spectra(x: np.ndarray, nperseg: int, step: int, count: int, win: np.ndarray, detrend: str,
                     nfft: int, n_half: int, psd_scale: Optional[float] = None) -> np.ndarray:
    """Rows ``x[i*step : i*step + nperseg]`` -> detrend (``apply_detrend``, spectral.rs:77-117) -> window -> zero-pad
    to the next power of two of ``nfft`` (what ``fft(&padded, None)`` does) -> ONE batched real-to-complex
    transform -> the first ``n_half`` bins, all on the device (``sfc_signal_spectra``: framing kernel, plan,
    strided copy back).  With ``psd_scale`` the device also sums |X|^2 over the rows and only ``n_half`` reals
    come back.  The reference keeps bins of the full complex transform; for real input those are the same numbers."""
    if detrend not in _DETREND:
        raise ValueError_(f"Unknown detrend option: {detrend}")
    lib = _lib.load()
    xs = np.ascontiguousarray(x, dtype=np.float64)
    w = np.ascontiguousarray(win, dtype=np.float64)
    if psd_scale is None:
        out = np.empty((count, n_half), dtype=np.complex128)
    else:
        out = np.empty(n_half, dtype=np.float64)
    check(lib.sfc_signal_spectra(_ptr(xs), xs.size, _ptr(w), nperseg, step, count, _next_pow2(max(nfft, 1)),
                                 _DETREND[detrend], 0 if psd_scale is None else 1, n_half,
                                 1.0 if psd_scale is None else float(psd_scale), _ptr(out)))
    return out


def periodogram(x, fs: Optional[float] = None, window: Optional[str] = None, nfft: Optional[int] = None,
                detrend: Optional[str] = None, scaling: Optional[str] = None) -> Tuple[np.ndarray, np.ndarray]:
    """spectral.rs:130-244."""
    a = np.asarray(x, dtype=np.float64).reshape(-1)
    if a.size == 0:
        raise ValueError_("Input array is empty")
    fs_val = 1.0 if fs is None else float(fs)
    nfft_val = a.size if nfft is None else int(nfft)
    window_val = "boxcar" if window is None else window
    detrend_val = "constant" if detrend is None else detrend
    scaling_val = "density" if scaling is None else scaling
    if fs_val <= 0.0:
        raise ValueError_(f"Sampling frequency must be positive, got {fs_val}")
    if nfft_val < a.size:
        raise ValueError_(f"NFFT must be at least as large as signal length, got {nfft_val} < {a.size}")
    win = spectral_window(window_val, a.size)
    scale = 1.0 / np.sum(win * win)
    n_half = nfft_val // 2 + nfft_val % 2
    spec = _segment_spectra(a, a.size, 1, 1, win, detrend_val, nfft_val, n_half)[0]
    psd = (spec.real ** 2 + spec.imag ** 2) * scale / (fs_val * a.size)
    if scaling_val != "density":
        psd = psd * fs_val
    return _fftfreq_head(nfft_val, fs_val, n_half), psd


def welch(x, fs: Optional[float] = None, window: Optional[str] = None, nperseg: Optional[int] = None,
          noverlap: Optional[int] = None, nfft: Optional[int] = None, detrend: Optional[str] = None,
          scaling: Optional[str] = None) -> Tuple[np.ndarray, np.ndarray]:
    """spectral.rs:257-410: all segments in one batched transform, then the average of |X|^2."""
    a = np.asarray(x, dtype=np.float64).reshape(-1)
    if a.size == 0:
        raise ValueError_("Input array is empty")
    fs_val = 1.0 if fs is None else float(fs)
    nperseg_val = min(256, a.size) if nperseg is None else int(nperseg)
    noverlap_val = nperseg_val // 2 if noverlap is None else int(noverlap)
    nfft_val = nperseg_val if nfft is None else int(nfft)
    window_val = "hann" if window is None else window
    detrend_val = "constant" if detrend is None else detrend
    scaling_val = "density" if scaling is None else scaling
    _check_common(fs_val, nfft_val, nperseg_val, noverlap_val)
    win = spectral_window(window_val, nperseg_val)
    scale = 1.0 / np.sum(win * win)
    step = nperseg_val - noverlap_val
    num_segments = (a.size - noverlap_val) // step if a.size >= noverlap_val else 0
    if num_segments < 1:
        raise ValueError_("Not enough data points for given nperseg and noverlap")
    n_half = nfft_val // 2 + nfft_val % 2
    # the reference divides by num_segments even when its loop breaks early on a short last segment
    usable = min(num_segments, (a.size - nperseg_val) // step + 1 if a.size >= nperseg_val else 0)
    psd = np.zeros(n_half)
    if usable > 0:
        psd = _segment_spectra(a, nperseg_val, step, usable, win, detrend_val, nfft_val, n_half,
                               psd_scale=scale / (fs_val * nperseg_val))
    psd /= num_segments
    if scaling_val != "density":
        psd = psd * fs_val
    return _fftfreq_head(nfft_val, fs_val, n_half), psd


def _apply_boundary(x: np.ndarray, nperseg: int, boundary: str) -> np.ndarray:
    """spectral.rs:413-447: nperseg/2 samples of zeros or of the edge value on both sides."""
    pad = nperseg // 2
    if boundary == "zeros":
        return np.concatenate([np.zeros(pad), x, np.zeros(pad)])
    if boundary == "extend":
        return np.concatenate([np.full(pad, x[0]), x, np.full(pad, x[-1])])
    if boundary == "none":
        return x
    raise ValueError_(f"Unknown boundary option: {boundary}")


def stft(x, fs: Optional[float] = None, window: Optional[str] = None, nperseg: Optional[int] = None,
         noverlap: Optional[int] = None, nfft: Optional[int] = None, detrend: Optional[str] = None,
         boundary: Optional[str] = None, padded: Optional[bool] = None
         ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """spectral.rs:468-628.  Returns (freqs, times, Z) with Z[segment][bin] (the reference's final
    transposition leaves segments as the outer index)."""
    a = np.asarray(x, dtype=np.float64).reshape(-1)
    if a.size == 0:
        raise ValueError_("Input array is empty")
    fs_val = 1.0 if fs is None else float(fs)
    nperseg_val = min(256, a.size) if nperseg is None else int(nperseg)
    noverlap_val = nperseg_val // 2 if noverlap is None else int(noverlap)
    nfft_val = nperseg_val if nfft is None else int(nfft)
    window_val = "hann" if window is None else window
    detrend_val = "constant" if detrend is None else detrend
    boundary_val = "zeros" if boundary is None else boundary
    padded_val = True if padded is None else bool(padded)
    _check_common(fs_val, nfft_val, nperseg_val, noverlap_val)
    win = spectral_window(window_val, nperseg_val)
    sig = _apply_boundary(a, nperseg_val, boundary_val) if padded_val else a
    step = nperseg_val - noverlap_val
    num_segments = (sig.size - noverlap_val) // step if sig.size >= noverlap_val else 0
    if num_segments < 1:
        raise ValueError_("Not enough data points for given nperseg and noverlap")
    n_half = nfft_val // 2 + nfft_val % 2
    times = (np.arange(num_segments) * step + nperseg_val // 2) / fs_val
    Z = np.zeros((num_segments, n_half), dtype=np.complex128)
    usable = min(num_segments, (sig.size - nperseg_val) // step + 1 if sig.size >= nperseg_val else 0)
    if usable > 0:
        Z[:usable] = _segment_spectra(sig, nperseg_val, step, usable, win, detrend_val, nfft_val, n_half)
    return _fftfreq_head(nfft_val, fs_val, n_half), times, Z


def spectrogram(x, fs: Optional[float] = None, window: Optional[str] = None, nperseg: Optional[int] = None,
                noverlap: Optional[int] = None, nfft: Optional[int] = None, detrend: Optional[str] = None,
                scaling: Optional[str] = None, mode: Optional[str] = None
                ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """spectral.rs:644-735 (modes psd / magnitude / angle / phase; "complex" is an error there too)."""
    mode_val = "psd" if mode is None else mode
    scaling_val = "density" if scaling is None else scaling
    n_in = np.asarray(x).size
    freqs, times, Z = stft(x, fs, window, nperseg, noverlap, nfft, detrend, "zeros", True)
    if mode_val == "psd":
        fs_val = 1.0 if fs is None else float(fs)
        nperseg_val = min(256, n_in) if nperseg is None else int(nperseg)
        win = spectral_window("hann" if window is None else window, nperseg_val)
        scale = 1.0 / np.sum(win * win)
        S = (Z.real ** 2 + Z.imag ** 2) * scale / (fs_val * nperseg_val)
        if scaling_val != "density":
            S = S * fs_val
    elif mode_val == "magnitude":
        S = np.abs(Z)
    elif mode_val in ("angle", "phase"):
        S = np.angle(Z)
    elif mode_val == "complex":
        raise ValueError_("Mode 'complex' returns complex values and is not supported for spectrogram")
    else:
        raise ValueError_(f"Unknown mode option: {mode_val}, expected 'psd', 'magnitude', 'angle', or 'phase'")
    return freqs, times, S


# ------------------------------------------------------------------------------------------------
# wiener.rs (the frequency-domain filters)
# ------------------------------------------------------------------------------------------------

@dataclass
class WienerConfig:
    """wiener.rs:50-84."""
    window_size: int = 15
    noise_power: Optional[float] = None
    frequency_domain: bool = True
    max_iterations: int = 1
    prior_snr: Optional[float] = None
    regularization: float = 1e-10
    boundary: bool = True


def _median_sorted(v: np.ndarray) -> float:
    n = v.size
    return float((v[n // 2 - 1] + v[n // 2]) / 2.0) if n % 2 == 0 else float(v[n // 2])


def estimate_noise_power(signal: np.ndarray) -> float:
    """wiener.rs:709-739: (1.4826 * median absolute deviation)^2."""
    v = np.sort(np.asarray(signal, dtype=np.float64).reshape(-1))
    med = _median_sorted(v)
    mad = _median_sorted(np.sort(np.abs(v - med)))
    return (1.4826 * mad) ** 2


def estimate_signal_power(signal: np.ndarray) -> float:
    """wiener.rs:742-750."""
    s = np.asarray(signal, dtype=np.float64).reshape(-1)
    return float(np.sum((s - s.mean()) ** 2) / s.size)


def wiener_filter_freq(signal, config: Optional[WienerConfig] = None) -> np.ndarray:
    """wiener.rs:137-196: gain = P / (P + snr * noise + reg) on every bin of the padded transform."""
    cfg = config or WienerConfig()
    s = np.asarray(signal, dtype=np.float64).reshape(-1)
    n = s.size
    noise_power = cfg.noise_power if cfg.noise_power is not None else estimate_noise_power(s)
    spec = fft(s, None)
    power = spec.real ** 2 + spec.imag ** 2
    snr = 1.0 if cfg.prior_snr is None else cfg.prior_snr
    gain = power / (power + snr * noise_power + cfg.regularization)
    return ifft(spec * gain, None)[:n].real.copy()


def wiener_filter(signal, noise_power: Optional[float] = None, window_size: Optional[int] = None) -> np.ndarray:
    """wiener.rs:110-127."""
    cfg = WienerConfig()
    if noise_power is not None:
        cfg.noise_power = noise_power
    if window_size is not None:
        cfg.window_size = window_size
    return wiener_filter_freq(signal, cfg)


def iterative_wiener_filter(signal, config: Optional[WienerConfig] = None) -> np.ndarray:
    """wiener.rs:304-345 (frequency-domain branch only; the time-domain filter has no FFT in it)."""
    cfg = config or WienerConfig()
    if not cfg.frequency_domain:
        raise ValueError_("only the frequency-domain Wiener filter is on the FFT path")
    s = np.asarray(signal, dtype=np.float64).reshape(-1)
    noise_power = cfg.noise_power if cfg.noise_power is not None else estimate_noise_power(s)
    cur = s.copy()
    for _ in range(cfg.max_iterations):
        sp = estimate_signal_power(cur)
        if sp < cfg.regularization:
            break
        it = WienerConfig(**{**cfg.__dict__, "prior_snr": sp / noise_power})
        cur = wiener_filter_freq(cur, it)
    return cur


def _mirror_gain(spec: np.ndarray, n: int, new_mag: np.ndarray) -> np.ndarray:
    """Bins 0..=n/2 get magnitude ``new_mag`` with their own phase; bin n-i gets the conjugate for
    0 < i < n/2 (wiener.rs:497-529 / 604-638: ``n`` is the SIGNAL length, the array may be longer)."""
    out = spec.copy()
    half = n // 2
    phase = np.angle(spec[: half + 1])
    out[: half + 1] = new_mag * np.exp(1j * phase)
    i = np.arange(1, half)
    if i.size:
        out[n - i] = new_mag[i] * np.exp(-1j * phase[i])
    return out


def spectral_subtraction(signal, noise_power: Optional[Sequence[float]] = None, alpha: Optional[float] = None,
                         beta: Optional[float] = None) -> np.ndarray:
    """wiener.rs:439-546."""
    s = np.asarray(signal, dtype=np.float64).reshape(-1)
    n = s.size
    a = 1.0 if alpha is None else alpha
    b = 0.01 if beta is None else beta
    spec = fft(s, None)
    if noise_power is not None:
        noise = np.asarray(noise_power, dtype=np.float64).reshape(-1)
    else:
        ns = int(min(n * 0.05, 100.0))
        if ns < 4:
            raise ValueError_("Signal too short to estimate noise spectrum")
        nf = fft(s[:ns], n)[: n // 2 + 1]
        noise = (nf.real ** 2 + nf.imag ** 2) / n
    half = n // 2
    idx = np.minimum(np.arange(half + 1), noise.size - 1)
    mag2 = np.abs(spec[: half + 1]) ** 2
    new_mag = np.sqrt(np.maximum(mag2 - a * noise[idx], b * mag2))
    return ifft(_mirror_gain(spec, n, new_mag), None)[:n].real.copy()


def smooth_psd(psd: np.ndarray) -> np.ndarray:
    """wiener.rs:775-794: moving average, window 2% of the length clamped to [3, 15]."""
    n = psd.size
    half = int(min(max(n * 0.02, 3.0), 15.0)) // 2
    c = np.concatenate([[0.0], np.cumsum(psd)])
    i = np.arange(n)
    lo, hi = np.maximum(i - half, 0), np.minimum(i + half + 1, n)
    return (c[hi] - c[lo]) / (hi - lo)


def psd_wiener_filter(signal, signal_psd: Optional[Sequence[float]] = None,
                      noise_psd: Optional[Sequence[float]] = None) -> np.ndarray:
    """wiener.rs:560-655."""
    s = np.asarray(signal, dtype=np.float64).reshape(-1)
    n = s.size
    spec = fft(s, None)
    half = n // 2
    if signal_psd is not None:
        s_psd = np.asarray(signal_psd, dtype=np.float64).reshape(-1)
    else:
        head = spec[: half + 1]
        s_psd = smooth_psd((head.real ** 2 + head.imag ** 2) / n)
    if noise_psd is not None:
        n_psd = np.asarray(noise_psd, dtype=np.float64).reshape(-1)
    else:
        n_psd = np.full(half + 1, estimate_noise_power(s))
    i = np.arange(half + 1)
    sp = np.where(i < s_psd.size, s_psd[np.minimum(i, s_psd.size - 1)], 0.0)
    npw = np.where(i < n_psd.size, n_psd[np.minimum(i, n_psd.size - 1)], 0.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        gain = np.where(sp + npw > 1e-10, sp / (sp + npw), 0.0)
    return ifft(_mirror_gain(spec, n, np.abs(spec[: half + 1]) * gain), None)[:n].real.copy()


# ------------------------------------------------------------------------------------------------
# window/mod.rs (the subset the callers below use) and streaming_stft.rs
# ------------------------------------------------------------------------------------------------

def signal_window(window_type: str, length: int, periodic: bool) -> np.ndarray:
    """window/mod.rs:38-83 for hann / hamming / blackman / bartlett / cosine / boxcar: periodic windows are the
    (length+1)-point symmetric window without its last sample (``_extend`` / ``_truncate``)."""
    if length == 0:
        raise ValueError_("Window length must be positive")
    if length <= 1:
        return np.ones(length)
    n = length + 1 if periodic else length
    i = np.arange(n, dtype=np.float64)
    w = window_type.lower()
    if w in ("hann", "hanning"):
        v = 0.5 * (1.0 - np.cos(2.0 * np.pi * i / (n - 1)))
    elif w == "hamming":
        v = 0.54 - 0.46 * np.cos(2.0 * np.pi * i / (n - 1))
    elif w == "blackman":
        v = 0.42 - 0.5 * np.cos(2.0 * np.pi * i / (n - 1)) + 0.08 * np.cos(4.0 * np.pi * i / (n - 1))
    elif w == "bartlett":
        m2 = (n - 1) / 2.0
        v = 1.0 - np.abs((i - m2) / m2)
    elif w == "cosine":
        v = np.sin(np.pi * i / (n - 1))
    elif w in ("boxcar", "rectangular"):
        v = np.ones(n)
    else:
        raise ValueError_(f"Unknown window type: {window_type}")
    return v[:length]


@dataclass
class StreamingStftConfig:
    """streaming_stft.rs:58-94."""
    frame_length: int = 512
    hop_length: int = 256
    window: str = "hann"
    center: bool = True
    pad_mode: str = "constant"
    magnitude_only: bool = False
    log_magnitude: bool = False
    power: float = 1.0
    log_epsilon: float = 1e-10


@dataclass
class StreamingStftStatistics:
    samples_processed: int
    frames_generated: int
    buffer_size: int
    latency_samples: int


class StreamingStft:
    """streaming_stft.rs:96-438.  ``process_batch`` gathers every frame the call completes and sends them to
    the device as one batched transform; the buffer bookkeeping is the reference's."""

    def __init__(self, config: Optional[StreamingStftConfig] = None):
        cfg = config or StreamingStftConfig()
        if cfg.frame_length == 0:
            raise ValueError_("Frame length must be greater than 0")
        if cfg.hop_length == 0:
            raise ValueError_("Hop length must be greater than 0")
        if cfg.hop_length > cfg.frame_length:
            raise ValueError_("Hop length should not exceed frame length")
        if cfg.power <= 0.0:
            raise ValueError_("Power must be positive")
        if cfg.center and cfg.pad_mode not in ("constant", "reflect", "symmetric"):
            raise ValueError_(f"Unknown pad mode: {cfg.pad_mode}")
        self.config = cfg
        self.window = signal_window(cfg.window, cfg.frame_length, True)
        self._buf: deque = deque()
        self.samples_processed = 0
        self.frames_generated = 0
        self._prefill()

    def _prefill(self) -> None:
        if self.config.center:  # every pad mode pre-fills zeros (streaming_stft.rs:166-190)
            self._buf.extend([0.0] * (self.config.frame_length // 2))

    # -- frame bookkeeping: returns the windowed frame this push completes, or None
    def _push(self, samples: np.ndarray) -> Optional[np.ndarray]:
        self._buf.extend(float(v) for v in samples)
        self.samples_processed += len(samples)
        L = self.config.frame_length
        if len(self._buf) < L:
            return None
        frame = np.fromiter(islice(self._buf, L), dtype=np.float64, count=L) * self.window
        for _ in range(min(self.config.hop_length, len(self._buf))):
            self._buf.popleft()
        self.frames_generated += 1
        return frame

    def _spectra(self, frames: List[np.ndarray]) -> List[np.ndarray]:
        """``compute_fft`` + ``process_spectrum`` (streaming_stft.rs:399-438) for a list of frames at once."""
        if not frames:
            return []
        L = self.config.frame_length
        P = _next_pow2(L)
        m = np.zeros((len(frames), P))
        m[:, :L] = np.stack(frames)
        spec = rfft_batch(m)[:, : L // 2 + 1]
        cfg = self.config
        if cfg.magnitude_only:
            mag = np.abs(spec)
            if cfg.power == 2.0:
                mag = spec.real ** 2 + spec.imag ** 2
            elif cfg.power != 1.0:
                mag = mag ** cfg.power
            if cfg.log_magnitude:
                mag = np.log(mag + cfg.log_epsilon)
            spec = mag.astype(np.complex128)
        return [row.copy() for row in spec]

    def process_frame(self, input_frame) -> Optional[np.ndarray]:
        f = self._push(np.asarray(input_frame, dtype=np.float64).reshape(-1))
        return None if f is None else self._spectra([f])[0]

    def process_batch(self, input_data, frame_size: int) -> List[np.ndarray]:
        d = np.asarray(input_data, dtype=np.float64).reshape(-1)
        frames = []
        start = 0
        while start + frame_size <= d.size:
            f = self._push(d[start:start + frame_size])
            if f is not None:
                frames.append(f)
            start += frame_size
        if start < d.size:
            f = self._push(d[start:])
            if f is not None:
                frames.append(f)
        return self._spectra(frames)

    def process_magnitude_frame(self, input_frame) -> Optional[np.ndarray]:
        s = self.process_frame(input_frame)
        if s is None:
            return None
        cfg = self.config
        mag = np.abs(s) if cfg.power == 1.0 else (s.real ** 2 + s.imag ** 2 if cfg.power == 2.0 else np.abs(s) ** cfg.power)
        return np.log(mag + cfg.log_epsilon) if cfg.log_magnitude else mag

    def get_latency_samples(self) -> int:
        c = self.config
        return c.frame_length // 2 + c.hop_length if c.center else c.frame_length

    def get_latency_seconds(self, sample_rate: float) -> float:
        return self.get_latency_samples() / sample_rate

    def get_statistics(self) -> StreamingStftStatistics:
        return StreamingStftStatistics(self.samples_processed, self.frames_generated, len(self._buf),
                                       self.get_latency_samples())

    def reset(self) -> None:
        self._buf.clear()
        self.samples_processed = 0
        self.frames_generated = 0
        self._prefill()

    def flush(self) -> List[np.ndarray]:
        L, hop = self.config.frame_length, self.config.hop_length
        frames = []
        while len(self._buf) >= hop:
            frame = np.zeros(L)
            avail = min(len(self._buf), L)
            frame[:avail] = list(islice(self._buf, avail))
            frames.append(frame * self.window)
            for _ in range(min(hop, len(self._buf))):
                self._buf.popleft()
            self.frames_generated += 1
        return self._spectra(frames)


@dataclass
class RealTimeStftStatistics:
    base_statistics: StreamingStftStatistics
    output_buffer_size: int
    output_buffer_capacity: int
    block_size: int


class RealTimeStft:
    """streaming_stft.rs:453-570: fixed-size input blocks, a bounded queue of finished spectra."""

    def __init__(self, config: Optional[StreamingStftConfig], block_size: int, max_buffer_size: int):
        self.streaming_stft = StreamingStft(config)
        self.block_size = int(block_size)
        self.max_output_buffer_size = int(max_buffer_size)
        self._out: deque = deque()

    def process_block(self, input_block) -> int:
        b = np.asarray(input_block, dtype=np.float64).reshape(-1)
        if b.size != self.block_size:
            raise ValueError_(f"Input block size {b.size} does not match expected size {self.block_size}")
        s = self.streaming_stft.process_frame(b)
        if s is None:
            return 0
        self._out.append(s)
        while len(self._out) > self.max_output_buffer_size:
            self._out.popleft()
        return 1

    def get_spectrum(self) -> Optional[np.ndarray]:
        return self._out.popleft() if self._out else None

    def get_all_spectra(self) -> List[np.ndarray]:
        r = list(self._out)
        self._out.clear()
        return r

    def peek_latest_spectrum(self) -> Optional[np.ndarray]:
        return self._out[-1] if self._out else None

    def available_spectra_count(self) -> int:
        return len(self._out)

    def is_buffer_full(self) -> bool:
        return len(self._out) >= self.max_output_buffer_size

    def reset(self) -> None:
        self.streaming_stft.reset()
        self._out.clear()

    def get_statistics(self) -> RealTimeStftStatistics:
        return RealTimeStftStatistics(self.streaming_stft.get_statistics(), len(self._out),
                                      self.max_output_buffer_size, self.block_size)


# ------------------------------------------------------------------------------------------------
# higher_order.rs (direct and Welch bispectrum, power spectrum)
# ------------------------------------------------------------------------------------------------

@dataclass
class HigherOrderConfig:
    """higher_order.rs:71-110 (``estimator``: "direct", "indirect" or "welch")."""
    estimator: str = "welch"
    fs: float = 1.0
    window: Optional[str] = "hann"
    n_segments: Optional[int] = None
    overlap: float = 0.5
    nfft: Optional[int] = None
    detrend: bool = True
    pad: bool = True
    non_redundant: bool = True


def _default_nfft(n: int) -> int:
    return max(2 ** int(np.ceil(np.log2(n))), 256)


def _direct_bispectra(spectra: np.ndarray, nfft: int) -> np.ndarray:
    """B[s][i][j] = X_s[i] X_s[j] conj(X_s[(i+j) % nfft]) on the full square (higher_order.rs:313-329: the
    triangle is mirrored, and the diagonal is written either way, so the square is symmetric and complete)."""
    nb = nfft // 2 + 1
    i = np.arange(nb)
    k = (i[:, None] + i[None, :]) % nfft
    head = spectra[:, :nb]
    return head[:, :, None] * head[:, None, :] * np.conj(spectra[:, k])


def _segment_rows(sig: np.ndarray, window: Optional[str], nfft: int, starts: Sequence[int], size: int) -> np.ndarray:
    """Windowed segments (each with a window of ITS OWN length: the last may be short) cut or zero-padded to
    ``nfft`` as ``fft(x, Some(nfft))`` does, transformed as one batch."""
    rows = np.zeros((len(starts), nfft), dtype=np.complex128)
    for r, st in enumerate(starts):
        seg = sig[st:min(st + size, sig.size)]
        if window is not None:
            seg = seg * signal_window(window, seg.size, True)
        m = min(seg.size, nfft)
        rows[r, :m] = seg[:m]
    return fftn(rows, None, [1]).reshape(len(starts), nfft)


def compute_bispectrum(signal, config: Optional[HigherOrderConfig] = None
                       ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """higher_order.rs:250-287 with the direct (:289-332) and Welch (:359-431) estimators."""
    cfg = config or HigherOrderConfig()
    s = np.asarray(signal, dtype=np.float64).reshape(-1)
    n = s.size
    if n < 4:
        raise ValueError_("Signal must have at least 4 data points")
    nfft = cfg.nfft if cfg.nfft is not None else _default_nfft(n)
    axis = np.linspace(0.0, cfg.fs / 2.0, nfft // 2 + 1)
    if cfg.estimator == "direct":
        B = _direct_bispectra(_segment_rows(s, cfg.window, nfft, [0], n), nfft)[0]
    elif cfg.estimator == "welch":
        size = min(nfft, n)
        ov = int(round(size * cfg.overlap))
        step = size - ov
        if step == 0:
            raise ValueError_("Overlap too large, resulting in zero step size")
        nseg = cfg.n_segments if cfg.n_segments is not None else int(np.floor((n - ov) / step))
        if nseg == 0:
            raise ValueError_("Signal too short for the specified segment size and overlap")
        starts = [i * step for i in range(nseg) if min(i * step + size, n) - i * step >= 4]
        nb = nfft // 2 + 1
        B = np.zeros((nb, nb), dtype=np.complex128)
        if starts:
            B = _direct_bispectra(_segment_rows(s, cfg.window, nfft, starts, size), nfft).sum(axis=0)
        B = B / nseg
    elif cfg.estimator == "indirect":
        B = compute_indirect_bispectrum(s, nfft, cfg.window)
    else:
        raise ValueError_(f"unknown bispectrum estimator {cfg.estimator!r}")
    return B, axis, axis.copy()


def bispectrum(signal, nfft: int, window: Optional[str] = None, n_segments: Optional[int] = None, fs: float = 1.0
               ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """higher_order.rs:142-165: magnitude of the Welch bispectrum."""
    B, f1, f2 = compute_bispectrum(signal, HigherOrderConfig(fs=fs, nfft=nfft, window=window, n_segments=n_segments))
    return np.abs(B), f1, f2


def compute_power_spectrum(signal, config: Optional[HigherOrderConfig] = None) -> Tuple[np.ndarray, np.ndarray]:
    """higher_order.rs:464-508: |X|^2 / nfft, doubled away from DC and Nyquist."""
    cfg = config or HigherOrderConfig()
    s = np.asarray(signal, dtype=np.float64).reshape(-1)
    nfft = cfg.nfft if cfg.nfft is not None else _default_nfft(s.size)
    nb = nfft // 2 + 1
    w = s * signal_window(cfg.window, s.size, True) if cfg.window is not None else s
    X = fft(w, nfft)[:nb]
    p = (X.real ** 2 + X.imag ** 2) / nfft
    if nb > 2:
        p[1:nb - 1] *= 2.0
    return p, np.linspace(0.0, cfg.fs / 2.0, nb)


# ------------------------------------------------------------------------------------------------
# cqt.rs
# ------------------------------------------------------------------------------------------------

@dataclass
class CqtConfig:
    """cqt.rs:24-68."""
    f_min: float = 32.7
    f_max: float = 8000.0
    bins_per_octave: int = 12
    q_factor: Optional[float] = None
    window_type: str = "hann"
    fs: float = 44100.0
    use_sparse: bool = True
    window_scaling: Optional[float] = None
    hop_size: Optional[int] = None


@dataclass
class SparseKernel:
    indices: np.ndarray
    values: np.ndarray
    normalization: float = 1.0


@dataclass
class CqtKernel:
    """cqt.rs:87-112.  ``spectra`` holds the kernel spectra as one [n_bins, n_fft] matrix with the entries
    the reference's sparse form drops (|K| <= 1e-6, cqt.rs:310-327) set to zero, so that applying the kernels
    is one matrix product; ``kernels`` gives the reference's per-bin sparse view of the same numbers."""
    spectra: np.ndarray
    mask: np.ndarray
    frequencies: np.ndarray
    n_fft: int
    q: float
    fs: float
    f_min: float
    f_max: float
    bins_per_octave: int

    @property
    def kernels(self) -> List[SparseKernel]:
        return [SparseKernel(np.nonzero(m)[0], row[m], 1.0) for row, m in zip(self.spectra, self.mask)]


@dataclass
class CqtResult:
    cqt: np.ndarray
    frequencies: np.ndarray
    kernel: Optional[CqtKernel] = None
    times: Optional[np.ndarray] = None


def _cqt_window(window_type: str, length: int) -> np.ndarray:
    """``create_window`` (cqt.rs:481-494): symmetric windows of window/mod.rs."""
    w = window_type.lower()
    if w not in ("hann", "hanning", "hamming", "blackman", "bartlett", "rectangular", "boxcar"):
        raise ValueError_(f"Unsupported window type: {window_type}")
    return signal_window(w, length, False)


def _odd_ceil(v: float) -> int:
    k = int(np.ceil(v))
    return k + 1 if k % 2 == 0 else k


def compute_cqt_kernel(f_min: float, f_max: float, bins_per_octave: int, q: float, fs: float, window_type: str,
                       window_scaling: Optional[float], use_sparse: bool) -> CqtKernel:
    """cqt.rs:229-347: windowed complex exponentials, unit energy, zero-padded to n_fft; all n_bins
    transforms run as one batched device FFT."""
    n_bins = int(np.ceil(np.log2(f_max / f_min) * bins_per_octave))
    freqs = f_min * 2.0 ** (np.arange(n_bins) / bins_per_octave)
    scale = 1.0 if window_scaling is None else window_scaling
    n_fft = _next_pow2(_odd_ceil(scale * q * fs / f_min))
    rows = np.zeros((n_bins, n_fft), dtype=np.complex128)
    for k, f in enumerate(freqs):
        klen = _odd_ceil(scale * q * fs / f)
        t = (np.arange(klen) - (klen - 1) / 2.0) / fs
        vals = np.exp(1j * (2.0 * np.pi * f * t)) * _cqt_window(window_type, klen)
        rows[k, :klen] = vals / np.sqrt(np.sum(vals.real ** 2 + vals.imag ** 2))
    spectra = fftn(rows, None, [1]).reshape(n_bins, n_fft) if n_bins else rows
    mask = np.abs(spectra) > 1e-6 if use_sparse else np.ones(spectra.shape, dtype=bool)
    return CqtKernel(np.where(mask, spectra, 0.0), mask, freqs, n_fft, q, fs, f_min, f_max, bins_per_octave)


def _cqt_rows(rows: np.ndarray, kernel: CqtKernel) -> np.ndarray:
    """[m, n_fft] real rows -> [m, n_bins]: sum over idx of FFT(row)[idx] * conj(K[bin][idx])."""
    X = fftn(rows.astype(np.complex128), None, [1]).reshape(rows.shape)
    return X @ np.conj(kernel.spectra).T


def compute_cqt_frame(signal, kernel: CqtKernel) -> np.ndarray:
    """cqt.rs:351-428: one padded transform, or the mean over consecutive n_fft chunks."""
    s = np.asarray(signal, dtype=np.float64).reshape(-1)
    n_fft = kernel.n_fft
    n_chunks = 1 if s.size < n_fft else int(np.ceil(s.size / n_fft))
    rows = np.zeros((n_chunks, n_fft))
    rows.reshape(-1)[: s.size] = s
    return _cqt_rows(rows, kernel).sum(axis=0) / n_chunks


def compute_cqt_spectrogram(signal, kernel: CqtKernel, hop_size: int) -> CqtResult:
    """cqt.rs:430-478: frame f = signal[f*hop : f*hop + n_fft], zero-padded; all frames in one batch."""
    s = np.asarray(signal, dtype=np.float64).reshape(-1)
    n_fft = kernel.n_fft
    n_frames = int(np.ceil(s.size / hop_size))
    start = np.arange(n_frames) * hop_size
    end = np.minimum(start + n_fft, s.size)
    times = (start + (end - start) // 2) / kernel.fs
    idx = start[:, None] + np.arange(n_fft)[None, :]
    rows = np.where(idx < s.size, s[np.minimum(idx, s.size - 1)], 0.0) if s.size else np.zeros((0, n_fft))
    cq = _cqt_rows(rows, kernel).T.copy() if n_frames else np.zeros((len(kernel.frequencies), 0), dtype=np.complex128)
    return CqtResult(cq, kernel.frequencies.copy(), kernel, times)


def constant_q_transform(signal, config: Optional[CqtConfig] = None) -> CqtResult:
    """cqt.rs:175-226."""
    cfg = config or CqtConfig()
    q = cfg.q_factor if cfg.q_factor is not None else 1.0 / (2.0 ** (1.0 / cfg.bins_per_octave) - 1.0)
    kernel = compute_cqt_kernel(cfg.f_min, cfg.f_max, cfg.bins_per_octave, q, cfg.fs, cfg.window_type,
                                cfg.window_scaling, cfg.use_sparse)
    if cfg.hop_size is not None:
        return compute_cqt_spectrogram(signal, kernel, cfg.hop_size)
    return CqtResult(compute_cqt_frame(signal, kernel).reshape(-1, 1), kernel.frequencies.copy(), kernel, None)


def cqt_magnitude(cqt: CqtResult, log_scale: bool = False, ref_value: Optional[float] = None) -> np.ndarray:
    """cqt.rs:515-540."""
    mag = np.abs(cqt.cqt)
    if log_scale:
        ref = ref_value if ref_value is not None else float(mag.max(initial=0.0))
        with np.errstate(divide="ignore"):
            mag = 20.0 * np.log10(mag / (ref + 1e-10))
    return mag


def cqt_phase(cqt: CqtResult) -> np.ndarray:
    """cqt.rs:551-563."""
    return np.angle(cqt.cqt)


def inverse_constant_q_transform(cqt: CqtResult, target_length: Optional[int] = None) -> np.ndarray:
    """cqt.rs:609-700: frame spectra = weighted sums of the kernel spectra, one batched inverse transform,
    overlap-add, peak normalisation."""
    kernel = cqt.kernel
    if kernel is None:
        raise ValueError_("CQT kernel not available for inverse transform")
    n_frames = cqt.cqt.shape[1]
    n_fft = kernel.n_fft
    if n_frames > 1 and cqt.times is not None:
        # f64::round: half away from zero (time differences are positive)
        hop = int(np.floor((cqt.times[1] - cqt.times[0]) * kernel.fs + 0.5)) if len(cqt.times) > 1 else n_fft // 2
    else:
        hop = n_fft
    out_len = target_length if target_length is not None else ((n_frames - 1) * hop + n_fft if n_frames > 1 else n_fft)
    out = np.zeros(out_len)
    if n_frames:
        spectra = cqt.cqt.T @ kernel.spectra  # [frames, n_fft]
        frames = ifftn(spectra, None, [1]).reshape(n_frames, n_fft).real
        for f in range(n_frames):
            st = f * hop
            en = min(st + n_fft, out_len)
            if en > st:
                out[st:en] += frames[f, : en - st]
    peak = float(np.max(np.abs(out), initial=0.0))
    return out / peak if peak > 0.0 else out


def _rem(a: int, b: int) -> int:
    """Rust's ``%`` on isize: the remainder takes the sign of the dividend."""
    return a - b * int(a / b)


def chromagram(cqt: CqtResult, n_chroma: Optional[int] = None, ref_note: Optional[int] = None) -> np.ndarray:
    """cqt.rs:713-760."""
    nc = 12 if n_chroma is None else n_chroma
    ref = (0 if ref_note is None else ref_note) % nc
    midi = 69.0 + 12.0 * np.log2(cqt.frequencies / 440.0)
    mag = np.abs(cqt.cqt)
    chroma = np.zeros((nc, cqt.cqt.shape[1]))
    for i, m in enumerate(midi):
        b = _rem(_rem(int(m), nc) + nc - ref, nc)
        if 0 <= b < nc:
            chroma[b] += mag[i]
    tot = chroma.sum(axis=0)
    return np.where(tot > 0.0, chroma / np.where(tot > 0.0, tot, 1.0), chroma)


# ------------------------------------------------------------------------------------------------
# higher_order.rs, continued: bicoherence and the indirect (triple-correlation) estimator
# ------------------------------------------------------------------------------------------------

def _round_half_away(v: np.ndarray) -> np.ndarray:
    """f64::round for non-negative values."""
    return np.floor(np.asarray(v, dtype=np.float64) + 0.5).astype(np.int64)


def compute_triple_correlation(signal, size: int) -> np.ndarray:
    """higher_order.rs:511-556: C(t1, t2) = sum_i c[i] c[i+t1] c[i+t2] / n for 0 <= t1, t2 < max_lag, stored in the
    lower-right quadrant of a (2 max_lag - 1)^2 matrix (the other quadrants stay zero there too)."""
    s = np.asarray(signal, dtype=np.float64).reshape(-1)
    n = s.size
    c = s - s.sum() / n
    max_lag = min(size, n // 3)
    if max_lag < 2:
        raise ValueError_("Signal too short for triple correlation calculation")
    # shifted[t][i] = c[i + t] (zero past the end): C = (c * shifted) @ shifted^T
    shifted = np.zeros((max_lag, n))
    for t in range(max_lag):
        shifted[t, : n - t] = c[t:]
    q = (shifted * c[None, :]) @ shifted.T / n
    out = np.zeros((2 * max_lag - 1, 2 * max_lag - 1))
    out[max_lag - 1:, max_lag - 1:] = q
    return out


def compute_2d_fft(matrix: np.ndarray, nfft: int) -> np.ndarray:
    """higher_order.rs:559-635: rows padded to max(cols, nfft) then to the next power of two (``fft(row, None)``), the
    first nfft/2+1 columns of that transformed the same way; two batched device passes."""
    m = np.asarray(matrix, dtype=np.float64)
    rows, cols = m.shape
    nb = nfft // 2 + 1
    p1 = _next_pow2(max(cols, nfft))
    a = np.zeros((rows, p1), dtype=np.complex128)
    a[:, :cols] = m
    r = fftn(a, None, [1]).reshape(rows, p1)[:, :nb]
    p2 = _next_pow2(max(rows, nfft))
    b = np.zeros((nb, p2), dtype=np.complex128)
    b[:, :rows] = r.T
    c = fftn(b, None, [1]).reshape(nb, p2)[:, :nb]
    return c.T.copy()


def compute_indirect_bispectrum(signal, nfft: int, window: Optional[str] = "hann") -> np.ndarray:
    """higher_order.rs:334-356."""
    s = np.asarray(signal, dtype=np.float64).reshape(-1)
    if window is not None:
        s = s * signal_window(window, s.size, True)
    return compute_2d_fft(compute_triple_correlation(s, nfft // 2 + 1), nfft)


def bicoherence(signal, nfft: int, window: Optional[str] = None, n_segments: Optional[int] = None, fs: float = 1.0
                ) -> Tuple[np.ndarray, Tuple[np.ndarray, np.ndarray]]:
    """higher_order.rs:192-247: |B(f1, f2)| / sqrt(P(f1) P(f2) P(f1 + f2)) with the Welch bispectrum."""
    cfg = HigherOrderConfig(fs=fs, nfft=nfft, window=window, n_segments=n_segments)
    B, f1, f2 = compute_bispectrum(signal, cfg)
    P, _ = compute_power_spectrum(signal, cfg)
    i_idx = _round_half_away(f1 * nfft / fs)
    j_idx = _round_half_away(f2 * nfft / fs)
    s_idx = _round_half_away((f1[:, None] + f2[None, :]) * nfft / fs) % nfft
    ok = (i_idx[:, None] < P.size) & (j_idx[None, :] < P.size) & (s_idx < P.size)
    ii = np.minimum(i_idx, P.size - 1)[:, None]
    jj = np.minimum(j_idx, P.size - 1)[None, :]
    norm = np.sqrt(P[ii] * P[jj] * P[np.minimum(s_idx, P.size - 1)])
    ok &= norm > 1e-10
    out = np.zeros(B.shape)
    out[ok] = np.abs(B)[ok] / norm[ok]
    return out, (f1, f2)


def _windowed_fft(signal, nfft: int, window: Optional[str]) -> np.ndarray:
    """``apply_window`` + ``compute_fft`` (higher_order.rs:433-461)."""
    s = np.asarray(signal, dtype=np.float64).reshape(-1)
    if window is not None:
        s = s * signal_window(window, s.size, True)
    return fft(s, nfft)


def trispectrum(signal, nfft: int, window: Optional[str] = None, fs: float = 1.0) -> np.ndarray:
    """higher_order.rs:638-684: the slice |X(f1) X(f2) X*(f1) X*(f2)|."""
    X = _windowed_fft(signal, nfft, window)[: nfft // 2 + 1]
    return np.abs((X[:, None] * X[None, :]) * np.conj(X)[:, None] * np.conj(X)[None, :])


def biamplitude(signal, nfft: int, window: Optional[str] = None, fs: float = 1.0
                ) -> Tuple[np.ndarray, Tuple[np.ndarray, np.ndarray]]:
    """higher_order.rs:698-745: |X(f1)| |X(f2)| |X(f1+f2)| where f1+f2 stays below Nyquist."""
    X = _windowed_fft(signal, nfft, window)
    nb = nfft // 2 + 1
    mag = np.abs(X)
    i = np.arange(nb)
    k = (i[:, None] + i[None, :]) % nfft
    out = np.where(k < nb, mag[:nb, None] * mag[None, :nb] * mag[np.minimum(k, nfft - 1)], 0.0)
    axis = np.linspace(0.0, fs / 2.0, nb)
    return out, (axis, axis.copy())


def cumulative_bispectrum(signal, nfft: int, window: Optional[str] = None, fs: float = 1.0
                          ) -> Tuple[np.ndarray, np.ndarray]:
    """higher_order.rs:762-804: mean bispectrum magnitude over the leading bw x bw squares, ten bandwidths."""
    mag, _, _ = bispectrum(signal, nfft, window, None, fs)
    nb = mag.shape[0]
    bandwidth = np.linspace(1.0, float(nb // 2), 10)
    out = np.zeros(10)
    for i, bw in enumerate(_round_half_away(bandwidth)):
        m = int(min(bw, nb))
        if bw > 0 and m > 0:
            out[i] = mag[:m, :m].sum() / (m * m)
    return out, bandwidth


def skewness_spectrum(signal, nfft: int, window: Optional[str] = None, fs: float = 1.0) -> Tuple[np.ndarray, np.ndarray]:
    """higher_order.rs:818-852: |B(f, f)| / P(f)^1.5."""
    cfg = HigherOrderConfig(fs=fs, nfft=nfft, window=window)
    B, f1, _ = compute_bispectrum(signal, cfg)
    P, _ = compute_power_spectrum(signal, cfg)
    nb = nfft // 2 + 1
    d = np.abs(np.diagonal(B))[:nb]
    ok = P[:nb] > 1e-10
    out = np.zeros(nb)
    out[ok] = d[ok] / P[:nb][ok] ** 1.5
    return out, f1


def detect_phase_coupling(signal, nfft: int, window: Optional[str] = None, fs: float = 1.0,
                          threshold: Optional[float] = None) -> List[Tuple[float, float, float]]:
    """higher_order.rs:868-912: local maxima (8 neighbours) of the bicoherence above ``threshold``, strongest first."""
    thresh = 0.5 if threshold is None else threshold
    b, (f1, f2) = bicoherence(signal, nfft, window, None, fs)
    peaks = []
    for i in range(1, b.shape[0] - 1):
        for j in range(1, b.shape[1] - 1):
            v = b[i, j]
            if v > thresh and v >= b[i - 1:i + 2, j - 1:j + 2].max():
                peaks.append((float(f1[i]), float(f2[j]), float(v)))
    peaks.sort(key=lambda p: -p[2])  # stable, like the reference's sort_by
    return peaks


# ------------------------------------------------------------------------------------------------
# hilbert.rs
# ------------------------------------------------------------------------------------------------

def hilbert(x) -> np.ndarray:
    """scirs2-signal hilbert.rs:58-150, as written: the spectrum comes from ``fft(x, None)`` (next power of two), its
    first n bins are multiplied by h (1 at DC and Nyquist, -2i on the positive half, 0 on the negative half) and an
    n-point inverse transform with 1/n follows.  (Not scirs2-fft's own ``hilbert``, which `consumers.py` mirrors.)"""
    s = np.asarray(x, dtype=np.float64).reshape(-1)
    n = s.size
    if n == 0:
        raise ValueError_("Input array is empty")
    spec = fft(s, None)[:n]
    h = np.ones(n, dtype=np.complex128)
    half = n // 2 if n % 2 == 0 else (n + 1) // 2
    h[1:half] = -2.0j
    h[half + 1 if n % 2 == 0 else half:] = 0.0
    return ifft(spec * h, n)


def envelope(x) -> np.ndarray:
    """hilbert.rs:180-196."""
    return np.abs(hilbert(x))


def _unwrap(phase: np.ndarray) -> np.ndarray:
    """hilbert.rs:258-279: differences folded into [-pi, pi] by repeated +-2 pi, then re-accumulated."""
    d = np.diff(phase)
    d = np.where(d > np.pi, d - 2.0 * np.pi * np.ceil((d - np.pi) / (2.0 * np.pi)), d)    # while d > pi: d -= 2 pi
    d = np.where(d < -np.pi, d + 2.0 * np.pi * np.ceil((-np.pi - d) / (2.0 * np.pi)), d)  # while d < -pi: d += 2 pi
    return np.concatenate([[phase[0]], phase[0] + np.cumsum(d)])


def instantaneous_phase(x, unwrap: bool = False) -> np.ndarray:
    """hilbert.rs:322-366."""
    a = hilbert(x)
    ph = np.arctan2(a.imag, a.real)
    return _unwrap(ph) if unwrap else ph


def instantaneous_frequency(x, fs: float) -> np.ndarray:
    """hilbert.rs:228-292: central differences of the unwrapped phase (one-sided at both ends)."""
    if np.asarray(x).size == 0:
        raise ValueError_("Input array is empty")
    if fs <= 0.0:
        raise ValueError_("Sampling frequency must be positive")
    u = instantaneous_phase(x, True)
    f = np.empty(u.size)
    f[0] = fs * (u[1] - u[0]) / (2.0 * np.pi)
    f[1:-1] = fs * (u[2:] - u[:-2]) / (4.0 * np.pi)
    f[-1] = fs * (u[-1] - u[-2]) / (2.0 * np.pi)
    return f


# ------------------------------------------------------------------------------------------------
# wvd.rs
# ------------------------------------------------------------------------------------------------

@dataclass
class WvdConfig:
    """wvd.rs:18-44."""
    analytic: bool = True
    time_window: Optional[np.ndarray] = None
    freq_window: Optional[np.ndarray] = None
    zero_padding: bool = True
    fs: float = 1.0


def _wvd_signal(signal, analytic: bool) -> np.ndarray:
    s = np.asarray(signal, dtype=np.float64).reshape(-1)
    return hilbert(s) if analytic else s.astype(np.complex128)


def compute_cross_wvd(s1: np.ndarray, s2: np.ndarray, config: WvdConfig) -> np.ndarray:
    """wvd.rs:232-344: for every time t the lag product s1[t+tau] conj(s2[t-tau]) (windowed) is laid out around
    n_fft/2 and transformed; all n transforms run as one batched device FFT.  Returns [n_fft/2+1, n]."""
    n = s1.size
    n_fft = 2 * n if config.zero_padding else n
    tw = None
    if config.time_window is not None:
        w = np.asarray(config.time_window, dtype=np.float64).reshape(-1)
        if w.size % 2 == 0:  # even length: a zero is inserted in the middle (wvd.rs:249-256)
            half = w.size // 2
            tw = np.zeros(w.size + 1)
            tw[:half] = w[:half]
            tw[half + 1:] = w[half:]
        else:
            tw = w.copy()
    fw = None
    if config.freq_window is not None:
        w = np.asarray(config.freq_window, dtype=np.float64).reshape(-1)
        if w.size < n_fft:
            fw = np.zeros(n_fft)
            off = (n_fft - w.size) // 2
            fw[off:off + w.size] = w
        elif w.size > n_fft:
            off = (w.size - n_fft) // 2
            fw = w[off:off + n_fft].copy()
        else:
            fw = w.copy()
    whl = tw.size // 2 if tw is not None else n // 2
    t = np.arange(n)[:, None]
    tau = np.arange(-whl, whl + 1)[None, :]
    valid = (tau >= -np.minimum(t, whl)) & (tau < np.minimum(n - t, whl + 1))
    i1, i2 = t + tau, t - tau
    valid &= (i1 >= 0) & (i1 < n) & (i2 >= 0) & (i2 < n)
    idx = tau + n_fft // 2
    valid &= (idx >= 0) & (idx < n_fft)
    vals = s1[np.clip(i1, 0, n - 1)] * np.conj(s2[np.clip(i2, 0, n - 1)])
    if tw is not None:
        wi = tau + whl
        vals = vals * np.where(wi < tw.size, tw[np.minimum(wi, tw.size - 1)], 0.0)
    P = _next_pow2(max(n_fft, 1))
    acorr = np.zeros((n, P), dtype=np.complex128)
    rows = np.broadcast_to(t, valid.shape)[valid]
    acorr[rows, np.broadcast_to(idx, valid.shape)[valid]] = vals[valid]
    if fw is not None:
        acorr[:, :n_fft] *= fw[None, :]
    spec = fftn(acorr, None, [1]).reshape(n, P)
    return spec[:, : n_fft // 2 + 1].T.copy()


def wigner_ville(signal, config: Optional[WvdConfig] = None) -> np.ndarray:
    """wvd.rs:79-90."""
    cfg = config or WvdConfig()
    a = _wvd_signal(signal, cfg.analytic)
    return compute_cross_wvd(a, a, cfg).real.copy()


def cross_wigner_ville(signal1, signal2, config: Optional[WvdConfig] = None) -> np.ndarray:
    """wvd.rs:124-154."""
    from .error import DimensionError

    cfg = config or WvdConfig()
    if np.asarray(signal1).size != np.asarray(signal2).size:
        raise DimensionError("Signals must have the same length for cross-WVD")
    return compute_cross_wvd(_wvd_signal(signal1, cfg.analytic), _wvd_signal(signal2, cfg.analytic), cfg)


def smoothed_pseudo_wigner_ville(signal, time_window, freq_window, config: Optional[WvdConfig] = None) -> np.ndarray:
    """wvd.rs:192-213."""
    base = config or WvdConfig()
    cfg = WvdConfig(base.analytic, np.asarray(time_window, dtype=np.float64), np.asarray(freq_window, dtype=np.float64),
                    base.zero_padding, base.fs)
    a = _wvd_signal(signal, cfg.analytic)
    return compute_cross_wvd(a, a, cfg).real.copy()


def frequency_axis(n_freqs: int, fs: float) -> np.ndarray:
    """wvd.rs:353-355."""
    return np.linspace(0.0, fs / 2.0, n_freqs)


def time_axis(n_times: int, fs: float) -> np.ndarray:
    """wvd.rs:367-370."""
    return np.linspace(0.0, (n_times - 1.0) / fs, n_times)
