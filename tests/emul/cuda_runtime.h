// TEST INFRASTRUCTURE ONLY — a minimal host stand-in for <cuda_runtime.h>, so that the tile kernel source
// (scirs_b200/csrc/fft_tile.cuh) and the planner (plan.cu) can be compiled by g++ with -DSFC_HOST_EMUL and the kernel's
// threads run as OS threads (one CTA at a time, __syncthreads() = a barrier).  "Device" memory is host memory.
// It exists to check INDEX LOGIC and planner plumbing on a machine without a GPU; it is never part of
// libscirs2_fft_cuda.so, nothing in the product includes it, and it says nothing about races, memory ordering or speed.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <algorithm>
#include <pthread.h>
#include <thread>
#include <vector>

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
[[noreturn]] inline void emul_unsupported(const char* what) {
    fprintf(stderr, "host emulation: %s not supported\n", what);
    abort();
}
using std::max;
using std::min;

// one launch: NT OS threads, each walking over all CTAs; the barrier after a CTA stands for its retirement
template <class K, class P>
inline void emul_launch(K kernel, const P& p, unsigned grid, int nt) {
    blockDim.x = (unsigned)nt;
    gridDim.x = grid;
    pthread_barrier_init(&emul_cta_barrier, nullptr, (unsigned)nt);
    std::vector<std::thread> th;
    th.reserve(nt);
    for (int t = 0; t < nt; ++t)
        th.emplace_back([&p, kernel, grid, t] {
            threadIdx.x = (unsigned)t;
            for (unsigned b = 0; b < grid; ++b) {
                blockIdx.x = b;
                kernel(p);
                pthread_barrier_wait(&emul_cta_barrier);
            }
        });
    for (auto& x : th) x.join();
    pthread_barrier_destroy(&emul_cta_barrier);
}

// ---- the slice of the runtime API plan.cu / kernel_inst.cuh use ----
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorUnknown = 999 };
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };

template <class T>
inline cudaError_t cudaMalloc(T** p, size_t bytes) {
    *p = (T*)aligned_alloc(256, (bytes + 255) / 256 * 256);
    return *p ? cudaSuccess : cudaErrorUnknown;
}
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "host emulation"; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
enum cudaStreamCaptureStatus { cudaStreamCaptureStatusNone = 0, cudaStreamCaptureStatusActive = 1 };
inline cudaError_t cudaStreamIsCapturing(cudaStream_t, cudaStreamCaptureStatus* s) { *s = cudaStreamCaptureStatusNone; return cudaSuccess; }
template <class F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 148; return cudaSuccess; }
template <class F>
inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 2; return cudaSuccess; }
