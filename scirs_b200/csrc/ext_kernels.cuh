// ext_kernels.cuh — the O(n) device passes around the FFT in the in-crate consumers of the hot path
// (SURVEY 8f rank 1): DCT/DST pre/post twiddles (scirs2-fft/src/dct.rs, dst.rs), Hartley and hfft
// output maps (hartley.rs:57-62, hfft/complex_to_real.rs:113-135), the Hilbert filter
// (lib.rs:470-510), the ihfft reflection (hfft/real_to_complex.rs:128-147) and STFT framing /
// output layout (spectrogram.rs:196-300).  All f64: the reference widens everything to f64 first.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sfc {

enum ElemMapMode : int32_t {
    EM_R2C_TAB = 0,    // dst(c)[o][j][i] = tab[j] * src(r)[o][j][i]
    EM_C2R_TAB = 1,    // dst(r)[o][k][i] = Re(tab[k] * src(c)[o][k][i])
    EM_C_LINCOMB = 2,  // dst(r)[e] = a * Re(src[e]) + b * Im(src[e])
    EM_C_TAB = 3,      // dst(c)[o][k][i] = src(c)[o][k][i] * tab[k]
    EM_IHFFT = 4,      // dst(c)[k] = k == 0 ? Re(src[0]) : (k < mid ? src[k] : conj(src[n - k])),  1-D, mid = (n+1)/2
};

struct MapParams {
    const void* src;
    void* dst;
    const void* tab;   // complex f64 table indexed along the axis (nullptr when unused)
    int64_t total;     // elements written
    int64_t n, inner;  // axis length and product of the trailing extents: j = (e / inner) % n
    int32_t mode;
    double a, b;
};
cudaError_t launch_map(const MapParams& p, cudaStream_t s);

// STFT framing: row f of dst (P reals) = (x_padded[f*step + j] - mean_f) * win[j], j < nperseg, then zeros
struct FrameParams {
    const double* x;     // the signal (len samples), boundary extension is index arithmetic
    const double* win;   // nperseg window samples
    double* dst;         // [frames][P]
    int64_t len, nperseg, step, frames, P;
    int32_t boundary;    // 0 none, 1 reflect, 2 zeros, 3 constant (spectrogram.rs:141-189), pad = nperseg each side
    int32_t detrend;     // 1: subtract the frame mean first (spectrogram.rs:253-257); 2: subtract the least-squares line
                         // (scirs2-signal spectral.rs:84-107)
};
cudaError_t launch_frames(const FrameParams& p, cudaStream_t s);

// STFT output: src [frames][src_pitch] complex -> dst [freq_len][frames], complex or a real-valued map of it
enum StftOut : int32_t { STFT_COMPLEX = 0, STFT_PSD = 1, STFT_MAGNITUDE = 2, STFT_PHASE = 3, STFT_ANGLE = 4 };
struct StftOutParams {
    const void* src;
    void* dst;
    int64_t frames, src_pitch, freq_len;
    int32_t mode;
    double scale;  // psd: |z|^2 * scale, magnitude: |z| * sqrt(scale)
};
cudaError_t launch_stft_out(const StftOutParams& p, cudaStream_t s);

// Welch reduction: out[k] = scale * sum over frames of |src[f][k]|^2, k < bins (src rows have src_pitch complex entries).
// Two deterministic stages: `parts` partial sums over contiguous frame ranges, then their sum.
struct PsdSumParams {
    const void* src;
    double* partial;  // [parts][bins]
    double* dst;      // [bins]
    int64_t frames, src_pitch, bins;
    int32_t parts;
    double scale;
};
cudaError_t launch_psd_sum(const PsdSumParams& p, cudaStream_t s);

}  // namespace sfc
