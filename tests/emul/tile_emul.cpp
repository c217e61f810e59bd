// TEST INFRASTRUCTURE ONLY — runs tile_fft_kernel instantiations on the host: one CTA at a time, its threads as OS
// threads, __syncthreads() as a pthread barrier.  Built by tests/test_kernel_emulation.py into tests/emul/_build/.
#define SFC_HOST_EMUL 1
#include <cuda_runtime.h>
#include <cstring>
#include <thread>
#include <vector>

thread_local emul_dim3 threadIdx;
thread_local emul_dim3 blockIdx;
emul_dim3 blockDim, gridDim;
pthread_barrier_t emul_cta_barrier;
namespace sfc {
alignas(128) unsigned char smem_raw[232448];  // the kernel's `extern __shared__` array: one CTA runs at a time
}

#include "fft_tile.cuh"

using namespace sfc;

template <typename T, int L, int TL, bool DBL, int E, int MODE>
static void run_grid(const PassParams& p, unsigned grid) {
    using C = TileCfg<T, L, TL, E, 1>;
    emul_launch(&tile_fft_kernel<T, L, TL, DBL, E, MODE, 1>, p, grid, C::NT);
}

// key = L * 1000000 + TL * 1000 + MODE (f64, E = 16)
extern "C" int emul_run(long long key, const PassParams* p, unsigned grid) {
#define CASE(L, TL, MODE) \
    case (long long)(L) * 1000000 + (TL) * 1000 + (MODE): run_grid<double, L, TL, false, 16, MODE>(*p, grid); return 0;
    switch (key) {
        CASE(512, 8, 1)
        CASE(512, 4, 1)
        CASE(256, 8, 1)
        CASE(64, 32, 7)
        CASE(512, 4, 7)
        CASE(256, 16, 7)
        CASE(512, 8, 8)
        CASE(64, 64, 5)
        default: return -1;
    }
}
extern "C" int emul_sizeof_params() { return (int)sizeof(PassParams); }
