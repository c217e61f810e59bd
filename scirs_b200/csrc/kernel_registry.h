// kernel_registry.h — table of compiled tile_fft_kernel instantiations.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include "pass_params.h"

namespace sfc {

enum Prec : int { PREC_F32 = 0, PREC_F64 = 1 };

struct KernelEntry {
    int prec;       // Prec
    int L;          // points per lane
    int TL;         // lanes per tile
    int E;          // points per thread (largest radix)
    int dbl;        // 1 = forward * table * inverse fused
    int mode;       // TileMode (0 generic, 1 fast c2c, 2 fast r2c, 3 fast c2r)
    int groups;     // independent thread groups per CTA (1 or 2)
    int threads;    // CTA size
    size_t smem;    // dynamic shared memory bytes
    const void* func;
    cudaError_t (*launch)(const PassParams& p, unsigned grid, cudaStream_t s);
};

// all entries (built once, thread-safe)
const KernelEntry* kernel_table(int* count);
const KernelEntry* find_kernel(int prec, int L, int TL, int dbl, int mode = 0);

// per-file registration hooks (one per kernels_*.cu translation unit)
void register_kernels_f64_small(void (*add)(const KernelEntry&));
void register_kernels_f64_mid(void (*add)(const KernelEntry&));
void register_kernels_f64_big(void (*add)(const KernelEntry&));
void register_kernels_f64_dbl_a(void (*add)(const KernelEntry&));
void register_kernels_f64_dbl_b(void (*add)(const KernelEntry&));
void register_kernels_f32_small(void (*add)(const KernelEntry&));
void register_kernels_f32_mid(void (*add)(const KernelEntry&));
void register_kernels_f32_big(void (*add)(const KernelEntry&));
void register_kernels_f32_dbl_a(void (*add)(const KernelEntry&));
void register_kernels_f32_dbl_b(void (*add)(const KernelEntry&));
void register_kernels_e8(void (*add)(const KernelEntry&));
void register_kernels_f64_real(void (*add)(const KernelEntry&));
void register_kernels_f32_real(void (*add)(const KernelEntry&));
void register_kernels_pipe(void (*add)(const KernelEntry&));
void register_kernels_dct(void (*add)(const KernelEntry&));
void register_kernels_pipe_dbl(void (*add)(const KernelEntry&));
void register_kernels_r3(void (*add)(const KernelEntry&));

}  // namespace sfc
