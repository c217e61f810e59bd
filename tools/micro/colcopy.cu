// Micro-benchmark: how fast can column tiles (SEG-byte row segments at a large stride) be moved at all?
// Same CTA shape as tile_fft_kernel column tiles: 256 threads, 16 x 16 B per thread, load all -> store all.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
struct P { const double2* in; double2* out; long stride; int rows; int tiles_per_batch; long batch_stride; int nbatch_fast; int tg; int smem_kb; };
template <int TL, int L, int MODE>
__global__ void __launch_bounds__(TL * L / 16, 2) colcopy(P p) {
    extern __shared__ double2 sm[];
    constexpr int NT = TL * L / 16, TPL = L / 16;
    const int tid = threadIdx.x, t = tid % TL, i = tid / TL;
    unsigned tile, batch;
    if (p.nbatch_fast) {
        unsigned per = p.nbatch_fast << p.tg, th = blockIdx.x / per, r = blockIdx.x - th * per;
        batch = r >> p.tg; tile = (th << p.tg) + (r & ((1u << p.tg) - 1));
    } else { tile = blockIdx.x % p.tiles_per_batch; batch = blockIdx.x / p.tiles_per_batch; }
    const long off = (long)batch * p.batch_stride + (long)tile * TL + t;
    double2 a[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) a[m] = p.in[off + (long)(i + m * TPL) * p.stride];
    if (MODE == 1) {  // emulate compute time: dependent FMA chain
#pragma unroll 1
        for (int k = 0; k < 40; ++k)
#pragma unroll
            for (int m = 0; m < 16; ++m) { a[m].x = fma(a[m].x, 1.0000001, a[m].y); a[m].y = fma(a[m].y, 0.9999999, a[m].x); }
    }
    sm[tid] = a[0]; __syncthreads(); a[0] = sm[(tid + 32) % NT];
#pragma unroll
    for (int m = 0; m < 16; ++m) p.out[off + (long)(i + m * TPL) * p.stride] = a[m];
}
template <int TL, int L, int MODE>
float run(P p, int grid, int smem) {
    cudaFuncSetAttribute(colcopy<TL, L, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 2; ++w) colcopy<TL, L, MODE><<<grid, TL * L / 16, smem>>>(p);
    cudaEventRecord(e0);
    for (int w = 0; w < 5; ++w) colcopy<TL, L, MODE><<<grid, TL * L / 16, smem>>>(p);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t e = cudaGetLastError(); if (e) printf("err %s\n", cudaGetErrorString(e));
    return ms / 5;
}
int main() {
    const long M = 1 << 21; const int NB = 16;  // 16 signals of 2^21 complex (537 MB) in, same out
    double2 *in, *out; cudaMalloc(&in, M * NB * 16); cudaMalloc(&out, M * NB * 16);
    cudaMemset(in, 0, M * NB * 16);
    const double gb = 2.0 * M * NB * 16 / 1e9;
    for (int smem_kb : {72, 36}) {
        for (int order = 0; order < 4; ++order) {
            P p{in, out, 0, 0, 0, M, 0, 0, smem_kb};
            if (order == 1) { p.nbatch_fast = NB; p.tg = 0; }
            if (order == 2) { p.nbatch_fast = NB; p.tg = 3; }
            if (order == 3) { p.nbatch_fast = NB; p.tg = 5; }
            // L1 = 1024 columns, stride L2 = 2048
            p.stride = 2048; p.tiles_per_batch = 2048 / 4;
            float a = run<4, 1024, 0>(p, p.tiles_per_batch * NB, smem_kb * 1024);
            float a1 = run<4, 1024, 1>(p, p.tiles_per_batch * NB, smem_kb * 1024);
            p.stride = 4096; p.tiles_per_batch = 4096 / 8;
            float b = run<8, 512, 0>(p, p.tiles_per_batch * NB, smem_kb * 1024);
            float b1 = run<8, 512, 1>(p, p.tiles_per_batch * NB, smem_kb * 1024);
            p.stride = 8192; p.tiles_per_batch = 8192 / 16;
            float c = run<16, 256, 0>(p, p.tiles_per_batch * NB, smem_kb * 1024);
            printf("smem %d KB order %d: 1024x4 (64 B) %.0f GB/s, +compute %.0f | 512x8 (128 B) %.0f, +compute %.0f | 256x16 (256 B) %.0f GB/s\n",
                   smem_kb, order, gb / a * 1e3, gb / a1 * 1e3, gb / b * 1e3, gb / b1 * 1e3, gb / c * 1e3);
        }
    }
    // rows for reference: stride 1 layout = TL rows of L contiguous
    return 0;
}
