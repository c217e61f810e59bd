// aux_kernels.cu — see aux_kernels.cuh
#include "aux_kernels.cuh"

namespace sfc {

template <typename TS, typename TD>
__global__ void nd_copy_kernel(const __grid_constant__ CopyParams p) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < p.total; idx += stride) {
        // dst multi-index (C order) -> src linear index, zero outside src
        int64_t rem = idx, sidx = 0, smul = 1;
        bool inb = true;
        for (int d = p.ndim - 1; d >= 0; --d) {
            const int64_t c = rem % p.dst_shape[d];
            rem /= p.dst_shape[d];
            inb = inb && (c < p.src_shape[d]);
            sidx += c * smul;
            smul *= p.src_shape[d];
        }
        double re = 0.0, im = 0.0;
        if (inb) {
            const TS* s = reinterpret_cast<const TS*>(p.src);
            if (p.src_complex) {
                re = (double)s[2 * sidx];
                im = (double)s[2 * sidx + 1];
            } else {
                re = (double)s[sidx];
            }
        }
        if (p.conj_src) im = -im;
        re *= p.scale;
        im *= p.scale;
        TD* d = reinterpret_cast<TD*>(p.dst);
        if (p.dst_complex) {
            d[2 * idx] = (TD)re;
            d[2 * idx + 1] = (TD)im;
        } else {
            d[idx] = (TD)re;
        }
    }
}

// the launch, or (tests/emul only, -DSFC_HOST_EMUL) the same kernel run on the host
#ifdef SFC_HOST_EMUL
#define SFC_AUX_LAUNCH(...) emul_launch(&__VA_ARGS__, p, (unsigned)blocks, threads)
#else
#define SFC_AUX_LAUNCH(...) __VA_ARGS__<<<(unsigned)blocks, threads, 0, s>>>(p)
#endif

cudaError_t launch_nd_copy(const CopyParams& p, cudaStream_t s) {
    if (p.total <= 0) return cudaSuccess;
    const int threads = 256;
    int64_t blocks = (p.total + threads - 1) / threads;
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (p.src_f64 && p.dst_f64)
        SFC_AUX_LAUNCH(nd_copy_kernel<double, double>);
    else if (!p.src_f64 && p.dst_f64)
        SFC_AUX_LAUNCH(nd_copy_kernel<float, double>);
    else if (p.src_f64 && !p.dst_f64)
        SFC_AUX_LAUNCH(nd_copy_kernel<double, float>);
    else
        SFC_AUX_LAUNCH(nd_copy_kernel<float, float>);
    return cudaGetLastError();
}

template <typename T>
__global__ void herm_fill_kernel(const __grid_constant__ HermParams p) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < p.total; idx += stride) {
        int64_t c[kMaxDims];
        int64_t rem = idx;
        for (int d = p.ndim - 1; d >= 0; --d) {
            c[d] = rem % p.out_shape[d];
            rem /= p.out_shape[d];
        }
        bool inb = true;
        for (int d = 0; d < p.ndim; ++d) inb = inb && (c[d] < p.x_shape[d]);
        bool cj = false;
        if (!inb) {
            // reflect through every listed axis, keeping DC and Nyquist (rfft.rs:861-879)
            int64_t r[kMaxDims];
            for (int d = 0; d < p.ndim; ++d) r[d] = c[d];
            for (int a = 0; a < p.naxes; ++a) {
                const int ax = p.axes[a];
                const int64_t n = p.out_shape[ax];
                if (c[ax] == 0 || (n % 2 == 0 && c[ax] == n / 2)) continue;
                r[ax] = n - c[ax];
            }
            inb = true;
            for (int d = 0; d < p.ndim; ++d) {
                c[d] = r[d];
                inb = inb && (c[d] < p.x_shape[d]);
            }
            cj = true;
        }
        T re = 0, im = 0;
        if (inb) {
            int64_t sidx = 0;
            for (int d = 0; d < p.ndim; ++d) sidx = sidx * p.x_shape[d] + c[d];
            const T* s = reinterpret_cast<const T*>(p.src);
            if (p.src_complex) {
                re = s[2 * sidx];
                im = s[2 * sidx + 1];
            } else {
                re = s[sidx];
            }
            if (cj) im = -im;
        }
        T* d = reinterpret_cast<T*>(p.dst);
        d[2 * idx] = re;
        d[2 * idx + 1] = im;
    }
}

cudaError_t launch_herm_fill(const HermParams& p, cudaStream_t s) {
    if (p.total <= 0) return cudaSuccess;
    const int threads = 256;
    int64_t blocks = (p.total + threads - 1) / threads;
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (p.f64)
        SFC_AUX_LAUNCH(herm_fill_kernel<double>);
    else
        SFC_AUX_LAUNCH(herm_fill_kernel<float>);
    return cudaGetLastError();
}

}  // namespace sfc
