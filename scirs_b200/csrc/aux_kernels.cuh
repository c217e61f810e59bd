// aux_kernels.cuh — the two non-FFT device passes of the path.
//
//  * nd_copy:    N-D pad / crop / real<->complex / f32->f64 convert / scale.
//                Reference: the element-wise convert + pad loops of fftn
//                (scirs2-fft/src/fft/algorithms.rs:617-664) and fft2 (:323-347).
//  * herm_fill:  Hermitian reconstruction of irfftn
//                (scirs2-fft/src/rfft.rs:733-901, reconstruct_hermitian_symmetry).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sfc {

constexpr int kMaxDims = 8;

struct CopyParams {
    int32_t ndim;
    int64_t dst_shape[kMaxDims];
    int64_t src_shape[kMaxDims];
    int32_t src_complex, dst_complex;  // 0 = real, 1 = complex
    int32_t src_f64, dst_f64;          // element precision
    int32_t conj_src;
    double scale;
    int64_t total;  // dst elements
    const void* src;
    void* dst;
};

struct HermParams {
    int32_t ndim;
    int64_t out_shape[kMaxDims];
    int64_t x_shape[kMaxDims];
    int32_t naxes;
    int32_t axes[kMaxDims];
    int32_t src_complex;  // 0: real input (imag = 0)
    int32_t f64;
    int64_t total;
    const void* src;
    void* dst;
};

cudaError_t launch_nd_copy(const CopyParams& p, cudaStream_t s);
cudaError_t launch_herm_fill(const HermParams& p, cudaStream_t s);

}  // namespace sfc
