// TEST INFRASTRUCTURE ONLY — a minimal host stand-in for <cuda_runtime.h>, so that the tile kernel source
// (scirs_b200/csrc/fft_tile.cuh) can be compiled by g++ and its threads run as OS threads (tile_emul.cpp).
// It exists to check the INDEX LOGIC of kernel flavours on a machine without a GPU; it is never part of
// libscirs2_fft_cuda.so and nothing in the product includes it.
#pragma once
#include <stdint.h>
#include <math.h>
#include <algorithm>
#include <pthread.h>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))
#define __shared__

struct emul_dim3 { unsigned x = 0, y = 0, z = 0; };
extern thread_local emul_dim3 threadIdx;
extern thread_local emul_dim3 blockIdx;
extern emul_dim3 blockDim, gridDim;
extern pthread_barrier_t emul_cta_barrier;

inline void __syncthreads() { pthread_barrier_wait(&emul_cta_barrier); }
inline void __syncwarp() {}
inline long long clock64() { return 0; }
inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
using std::max;
using std::min;
