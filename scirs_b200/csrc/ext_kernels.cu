// ext_kernels.cu — see ext_kernels.cuh.  Plain HBM-bound element-wise passes: 128-bit accesses,
// grid-stride loops sized to the SM count.
#include "ext_kernels.cuh"

namespace sfc {

namespace {

__device__ __forceinline__ double2 cmul2(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}

__global__ void __launch_bounds__(256) map_kernel(const MapParams p) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const double2* __restrict__ tab = reinterpret_cast<const double2*>(p.tab);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < p.total; e += stride) {
        const int64_t j = (e / p.inner) % p.n;
        switch (p.mode) {
            case EM_R2C_TAB: {
                const double x = reinterpret_cast<const double*>(p.src)[e];
                const double2 t = tab[j];
                reinterpret_cast<double2*>(p.dst)[e] = make_double2(t.x * x, t.y * x);
                break;
            }
            case EM_C2R_TAB: {
                const double2 z = reinterpret_cast<const double2*>(p.src)[e];
                const double2 t = tab[j];
                reinterpret_cast<double*>(p.dst)[e] = fma(t.x, z.x, -(t.y * z.y));
                break;
            }
            case EM_C_LINCOMB: {
                const double2 z = reinterpret_cast<const double2*>(p.src)[e];
                reinterpret_cast<double*>(p.dst)[e] = p.a * z.x + p.b * z.y;
                break;
            }
            case EM_C_TAB: {
                const double2 z = reinterpret_cast<const double2*>(p.src)[e];
                reinterpret_cast<double2*>(p.dst)[e] = cmul2(z, tab[j]);
                break;
            }
            default: {  // EM_IHFFT
                const double2* z = reinterpret_cast<const double2*>(p.src);
                const int64_t mid = (p.n + 1) / 2;
                double2 v;
                if (e == 0) v = make_double2(z[0].x, 0.0);
                else if (e < mid) v = z[e];
                else {
                    v = z[p.n - e];
                    v.y = -v.y;
                }
                reinterpret_cast<double2*>(p.dst)[e] = v;
            }
        }
    }
}

// one CTA per frame (grid-stride over frames)
__device__ __forceinline__ double padded_sample(const FrameParams& p, int64_t t) {
    if (p.boundary == 0) return p.x[t];
    const int64_t pad = p.nperseg;
    if (t < pad) {
        if (p.boundary == 1) return p.x[pad - 1 - t];
        return p.boundary == 2 ? 0.0 : p.x[0];
    }
    if (t < pad + p.len) return p.x[t - pad];
    if (p.boundary == 1) return p.x[p.len - 1 - (t - pad - p.len)];
    return p.boundary == 2 ? 0.0 : p.x[p.len - 1];
}

// sum of `v` over the CTA (256 threads), returned to every thread
__device__ __forceinline__ double block_sum(double v, double* red, double* bcast) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) *bcast = t;
    }
    __syncthreads();
    const double r = *bcast;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(256) frame_kernel(const FrameParams p) {
    __shared__ double red[32];
    __shared__ double bc;
    for (int64_t f = blockIdx.x; f < p.frames; f += gridDim.x) {
        const int64_t start = f * p.step;
        double mean = 0.0, slope = 0.0;
        if (p.detrend) {
            double s = 0.0, sj = 0.0;
            for (int64_t j = threadIdx.x; j < p.nperseg; j += blockDim.x) {
                const double v = padded_sample(p, start + j);
                s += v;
                sj += (double)j * v;
            }
            const double sum_y = block_sum(s, red, &bc);
            const double n = (double)p.nperseg;
            mean = sum_y / n;
            if (p.detrend == 2) {
                // y - (slope * j + intercept), the regression sums of spectral.rs:88-100 in closed form
                const double sum_xy = block_sum(sj, red, &bc);
                const double sum_x = n * (n - 1.0) * 0.5, sum_xx = (n - 1.0) * n * (2.0 * n - 1.0) / 6.0;
                slope = (n * sum_xy - sum_x * sum_y) / (n * sum_xx - sum_x * sum_x);
                mean = (sum_y - slope * sum_x) / n;  // the intercept
            }
        }
        double* row = p.dst + f * p.P;
        for (int64_t j = threadIdx.x; j < p.P; j += blockDim.x)
            row[j] = j < p.nperseg ? (padded_sample(p, start + j) - (slope * (double)j + mean)) * p.win[j] : 0.0;
    }
}

// one thread per bin, consecutive threads on consecutive bins (coalesced 16-byte loads); blockIdx.y = frame range
__global__ void __launch_bounds__(256) psd_partial_kernel(const PsdSumParams p) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= p.bins) return;
    const int64_t per = (p.frames + p.parts - 1) / p.parts;
    const int64_t f0 = (int64_t)blockIdx.y * per;
    const int64_t f1 = f0 + per < p.frames ? f0 + per : p.frames;
    const double2* __restrict__ z = reinterpret_cast<const double2*>(p.src);
    double acc = 0.0;
    for (int64_t f = f0; f < f1; ++f) {
        const double2 v = z[f * p.src_pitch + k];
        acc += v.x * v.x + v.y * v.y;
    }
    p.partial[(int64_t)blockIdx.y * p.bins + k] = acc;
}

__global__ void __launch_bounds__(256) psd_final_kernel(const PsdSumParams p) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= p.bins) return;
    double acc = 0.0;
    for (int q = 0; q < p.parts; ++q) acc += p.partial[(int64_t)q * p.bins + k];
    p.dst[k] = acc * p.scale;
}

// 32 x 32 tile transpose through shared memory: reads rows of frames, writes rows of frequencies
__global__ void __launch_bounds__(256) stft_out_kernel(const StftOutParams p) {
    __shared__ double2 tile[32][33];
    const int64_t f0 = (int64_t)blockIdx.x * 32, k0 = (int64_t)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const double2* __restrict__ src = reinterpret_cast<const double2*>(p.src);
    for (int r = ty; r < 32; r += 8) {
        const int64_t f = f0 + r, k = k0 + tx;
        tile[r][tx] = (f < p.frames && k < p.freq_len) ? src[f * p.src_pitch + k] : make_double2(0.0, 0.0);
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int64_t k = k0 + r, f = f0 + tx;
        if (k >= p.freq_len || f >= p.frames) continue;
        const double2 z = tile[tx][r];
        const int64_t o = k * p.frames + f;
        switch (p.mode) {
            case STFT_COMPLEX: reinterpret_cast<double2*>(p.dst)[o] = z; break;
            case STFT_PSD: reinterpret_cast<double*>(p.dst)[o] = (z.x * z.x + z.y * z.y) * p.scale; break;
            case STFT_MAGNITUDE: reinterpret_cast<double*>(p.dst)[o] = hypot(z.x, z.y) * sqrt(p.scale); break;
            case STFT_PHASE: reinterpret_cast<double*>(p.dst)[o] = atan2(z.y, z.x); break;
            default: reinterpret_cast<double*>(p.dst)[o] = atan2(z.y, z.x) * (180.0 / 3.14159265358979323846); break;
        }
    }
}

int grid_for(int64_t work, int per_block) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = (work + per_block - 1) / per_block;
    const int64_t cap = (int64_t)sms * 8;  // a multiple of the SM count, 8 resident CTAs of 256 threads each
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace

cudaError_t launch_map(const MapParams& p, cudaStream_t s) {
    if (p.total <= 0) return cudaSuccess;
    map_kernel<<<grid_for(p.total, 256), 256, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_frames(const FrameParams& p, cudaStream_t s) {
    if (p.frames <= 0) return cudaSuccess;
    frame_kernel<<<grid_for(p.frames, 1), 256, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_psd_sum(const PsdSumParams& p, cudaStream_t s) {
    if (p.bins <= 0) return cudaSuccess;
    dim3 grid((unsigned)((p.bins + 255) / 256), (unsigned)p.parts);
    psd_partial_kernel<<<grid, 256, 0, s>>>(p);
    psd_final_kernel<<<grid.x, 256, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_stft_out(const StftOutParams& p, cudaStream_t s) {
    if (p.frames <= 0 || p.freq_len <= 0) return cudaSuccess;
    dim3 grid((unsigned)((p.frames + 31) / 32), (unsigned)((p.freq_len + 31) / 32));
    stft_out_kernel<<<grid, 256, 0, s>>>(p);
    return cudaGetLastError();
}

}  // namespace sfc
