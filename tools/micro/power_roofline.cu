// power_roofline.cu — what bandwidth can ANY kernel sustain on this part when it moves HBM-rate data AND executes the FP64
// / shared-memory work of an FFT tile?  A streaming kernel with the tile kernel's shape (256 threads, 16 complex points per
// thread, 16 B loads / stores at a 256-element stride) does K FP64 instructions per point on independent chains and S
// shared-memory exchanges (STS.128, barrier, LDS.128 of another thread's slot), nothing else.  The 4096-point c2c tile
// executes 694 FP64 instructions per 16 points = 43 per point and 2 exchanges.  Sustained: 1.5 s preload, >= 1 s timed,
// SM clock and power sampled over NVML meanwhile.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o power_roofline power_roofline.cu -ldl
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <chrono>
#include <thread>
#include <atomic>
#include <dlfcn.h>
#include <cuda_runtime.h>

template <int K, int S, bool ADD, int NT = 256, int MINB = 2, bool PERSIST = false, int MODE = 0, int PF = 0>
__global__ void __launch_bounds__(NT, MINB) stream_kernel(const double2* __restrict__ in, double2* __restrict__ out, double a, double b,
                                                          unsigned ntiles) {
    extern __shared__ __align__(128) double2 sm[];
  for (unsigned tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const size_t base = (size_t)tile * (16 * NT) + threadIdx.x;
    if (PF > 0 && threadIdx.x == 0 && tile + PF < ntiles)  // pull the tile some CTA will need PF tiles from now into L2
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(in + (size_t)(tile + PF) * (16 * NT)), "r"(16 * NT * 16) : "memory");
    double2 v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = in[base + m * NT];
#pragma unroll
    for (int s = 0; s <= S; ++s) {
        constexpr int KS = K / (S + 1);
#pragma unroll
        for (int k = 0; k < KS / 2; ++k) {
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                if (ADD) { v[m].x = v[m].x + b; v[m].y = v[m].y + a; }
                else { v[m].x = fma(v[m].x, a, b); v[m].y = fma(v[m].y, a, b); }
            }
        }
        if (s < S && MODE == 5 && s == 1) {
            // second exchange inside the warp
            double2* w = sm + (threadIdx.x >> 5) * (17 * 32);
            const int l = threadIdx.x & 31;
#pragma unroll
            for (int m = 0; m < 16; ++m) w[17 * l + m] = v[m];
            __syncwarp();
#pragma unroll
            for (int m = 0; m < 16; ++m) v[m] = w[l + 34 * m];
            __syncwarp();
        } else if (s < S && MODE == 1) {
            // warp-local exchange: same traffic, no CTA-wide barrier (a warp's 512 slots are its own)
            double2* w = sm + (threadIdx.x >> 5) * (17 * 32);
            const int l = threadIdx.x & 31;
#pragma unroll
            for (int m = 0; m < 16; ++m) w[17 * l + m] = v[m];
            __syncwarp();
#pragma unroll
            for (int m = 0; m < 16; ++m) v[m] = w[l + 34 * m];
            __syncwarp();
        } else if (s < S && MODE == 2) {
            // half the data: real parts only, 8-byte slots
            double* h = reinterpret_cast<double*>(sm);
#pragma unroll
            for (int m = 0; m < 16; ++m) h[17 * threadIdx.x + m] = v[m].x;
            __syncthreads();
#pragma unroll
            for (int m = 0; m < 16; ++m) v[m].x = h[threadIdx.x + (NT + NT / 16) * m];
            __syncthreads();
        } else if (s < S && MODE == 3) {
            if (a == 123.456) sm[threadIdx.x] = v[0];  // shared memory allocated, never touched
        } else if (s < S) {
            // padded exchange like the tile kernel's: thread t writes slots 17 t + m, reads t + 272 m (another thread's data)
#pragma unroll
            for (int m = 0; m < 16; ++m) sm[17 * threadIdx.x + m] = v[m];
            __syncthreads();
#pragma unroll
            for (int m = 0; m < 16; ++m) v[m] = sm[threadIdx.x + (NT + NT / 16) * m];
            __syncthreads();
        }
    }
    if (MODE == 5) {
        // thread (k0, k1): k0 = (t & 1) + 2 (t >> 5), k1 = (t >> 1) & 15 -> element k0 + 16 k1 + 256 m: 32-byte segments
        const int t = threadIdx.x;
        const size_t ob = (size_t)tile * (16 * NT) + ((t & 1) + 2 * (t >> 5)) + 16 * ((t >> 1) & 15);
#pragma unroll
        for (int m = 0; m < 16; ++m) out[ob + m * NT] = v[m];
    } else {
#pragma unroll
        for (int m = 0; m < 16; ++m) out[base + m * NT] = v[m];
    }
    if (!PERSIST) break;
  }
}

// random payload: constant data toggles no wires and draws far less power (first cut: zeros copied at 6.8 TB/s for 761 W)
__global__ void fill_random(double2* p, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned long long z = i * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
        unsigned long long w = z * 0xD6E8FEB86659FD93ull; w ^= w >> 32;
        p[i] = make_double2((double)(long long)z * (1.0 / 9223372036854775808.0), (double)(long long)w * (1.0 / 9223372036854775808.0));
    }
}


// Candidate structure for the 4096-point tile: the row is LANDED in shared memory by one bulk copy (no LSU global loads, no
// registers), every warp transforms a 512-slot region of its own through two rounds (read 16, compute, write back in place,
// __syncwarp) — the first two radix-16 stages with the first exchange inside a warp — and only the last exchange synchronises
// the CTA.  Shared-memory traffic: 3 reads + 2 writes of the tile (5 x 64 KiB) against 2 + 2 today plus the global loads that
// no longer pass through the LSU.  Access patterns are conflict-free stand-ins (slot bijections), not the real index algebra.
template <int K, bool ADD>
__global__ void __launch_bounds__(256, 2) landed_kernel(const double2* __restrict__ in, double2* __restrict__ out, double a, double b,
                                                        unsigned ntiles) {
    extern __shared__ __align__(128) double2 sm[];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned tile = blockIdx.x;
    const unsigned sbar = (unsigned)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sbar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sbar), "r"(65536) : "memory");
        const unsigned dst = (unsigned)__cvta_generic_to_shared(sm);
#pragma unroll
        for (int c = 0; c < 4; ++c)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + c * 16384),
                         "l"(in + (size_t)tile * 4096 + c * 1024), "r"(16384), "r"(sbar)
                         : "memory");
    }
    {  // everybody waits for the landing
        unsigned done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(sbar), "r"(0) : "memory");
    }
    const int l = threadIdx.x & 31;
    double2* w = sm + (threadIdx.x >> 5) * 512;  // the warp's own 512 slots
    double2 v[16];
    constexpr int KS = K / 3;
#pragma unroll
    for (int round = 0; round < 2; ++round) {
        // read 16 slots of the warp's region (lane + 32 m: conflict-free), compute, write back to the same slots
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = w[((l + 32 * m) + 7 * round * m) & 511];
#pragma unroll
        for (int k = 0; k < KS / 2; ++k)
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                if (ADD) { v[m].x = v[m].x + b; v[m].y = v[m].y + a; }
                else { v[m].x = fma(v[m].x, a, b); v[m].y = fma(v[m].y, a, b); }
            }
#pragma unroll
        for (int m = 0; m < 16; ++m) w[((l + 32 * m) + 7 * round * m) & 511] = v[m];
        if (round == 0) __syncwarp();
    }
    __syncthreads();
    // the one CTA-wide exchange: thread t reads slots t + 256 m (other warps' regions), last stage, coalesced stores
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = sm[threadIdx.x + 256 * m];
#pragma unroll
    for (int k = 0; k < KS / 2; ++k)
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            if (ADD) { v[m].x = v[m].x + b; v[m].y = v[m].y + a; }
            else { v[m].x = fma(v[m].x, a, b); v[m].y = fma(v[m].y, a, b); }
        }
    const size_t base = (size_t)tile * 4096 + threadIdx.x;
#pragma unroll
    for (int m = 0; m < 16; ++m) out[base + m * 256] = v[m];
}

typedef int (*nvml_fn)(...);
struct Nvml {
    void* h = nullptr; void* dev = nullptr;
    int (*init)() = nullptr; int (*get)(unsigned, void**) = nullptr; int (*clk)(void*, int, unsigned*) = nullptr; int (*pwr)(void*, unsigned*) = nullptr;
    bool ok = false;
    Nvml() {
        h = dlopen("libnvidia-ml.so.1", RTLD_NOW);
        if (!h) return;
        init = (int (*)())dlsym(h, "nvmlInit_v2");
        get = (int (*)(unsigned, void**))dlsym(h, "nvmlDeviceGetHandleByIndex_v2");
        clk = (int (*)(void*, int, unsigned*))dlsym(h, "nvmlDeviceGetClockInfo");
        pwr = (int (*)(void*, unsigned*))dlsym(h, "nvmlDeviceGetPowerUsage");
        ok = init && get && clk && pwr && init() == 0 && get(0, &dev) == 0;
    }
};

template <typename F>
void run_launch(const char* name, F launch, size_t n, Nvml& nv);

template <int K, bool ADD>
void run_landed(const char* name, const double2* in, double2* out, size_t n, Nvml& nv, int smem_kb = 64) {
    const unsigned ntiles = (unsigned)(n / 4096);
    cudaFuncSetAttribute(landed_kernel<K, ADD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
    run_launch(name, [&]() { landed_kernel<K, ADD><<<ntiles, 256, smem_kb * 1024>>>(in, out, 0.9173, 0.3391, ntiles); }, n, nv);
}

template <int K, int S, bool ADD, int NT = 256, int MINB = 2, bool PERSIST = false, int MODE = 0, int PF = 0>
void run(const char* name, const double2* in, double2* out, size_t n, Nvml& nv) {
    const unsigned ntiles = (unsigned)(n / (16 * NT));
    const unsigned grid = PERSIST ? 148 * MINB : ntiles;
    const size_t smem = S ? (size_t)(17 * NT + 16) * 16 : 0;
    cudaFuncSetAttribute(stream_kernel<K, S, ADD, NT, MINB, PERSIST, MODE, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    run_launch(name, [&]() { stream_kernel<K, S, ADD, NT, MINB, PERSIST, MODE, PF><<<grid, NT, smem>>>(in, out, 0.9173, 0.3391, ntiles); }, n, nv);
}

template <typename F>
void run_launch(const char* name, F launch, size_t n, Nvml& nv) {
    auto t0 = std::chrono::steady_clock::now();
    while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < 1.5) {
        for (int i = 0; i < 16; ++i) launch();
        cudaDeviceSynchronize();
    }
    std::atomic<bool> stop{false};
    std::vector<unsigned> clk, pw;
    std::thread th([&]() {
        while (!stop && nv.ok) {
            unsigned c = 0, p = 0;
            if (nv.clk(nv.dev, 1 /* SM */, &c) == 0 && nv.pwr(nv.dev, &p) == 0) { clk.push_back(c); pw.push_back(p); }
            std::this_thread::sleep_for(std::chrono::milliseconds(50));
        }
    });
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 800;
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    stop = true; th.join();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    std::sort(clk.begin(), clk.end()); std::sort(pw.begin(), pw.end());
    const double gbs = 2.0 * n * 16 / ms / 1e6;
    printf("%-44s %7.4f ms  %7.1f GB/s  %5.1f%% of 6553.9   sm %4u MHz  %4u W   (%s)\n", name, ms, gbs, gbs / 65.539,
           clk.empty() ? 0 : clk[clk.size() / 2], pw.empty() ? 0 : pw[pw.size() / 2] / 1000, cudaGetErrorString(cudaGetLastError()));
    fflush(stdout);
}

int main() {
    const size_t n = (size_t)65536 * 4096;  // complex points: 4 GiB in, 4 GiB out (the headline batch)
    double2 *in, *out;
    if (cudaMalloc(&in, n * 16) != cudaSuccess || cudaMalloc(&out, n * 16) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    fill_random<<<148 * 8, 256>>>(in, n);
    cudaDeviceSynchronize();
    Nvml nv;
    printf("streaming kernel, 65,536 x 4096 complex f64 (8.59 GB per launch), random payload, sustained; NVML %s\n", nv.ok ? "ok" : "unavailable");
    run<0, 0, false>("copy (0 FP64 / point, no exchange)", in, out, n, nv);
    run<42, 2, false>("42 DFMA / point, 2 exchanges (the 4096 tile)", in, out, n, nv);
    run<42, 2, false, 256, 2, false, 1>("42 DFMA, 2 warp-local exchanges", in, out, n, nv);
    run<42, 2, false, 256, 2, false, 5>("42 DFMA, CTA exchange + warp-local exchange + 32 B store segments", in, out, n, nv);
    run<42, 2, true, 256, 2, false, 5>("42 DADD, CTA exchange + warp-local exchange + 32 B store segments", in, out, n, nv);
    run<42, 2, true>("42 DADD / point, 2 exchanges", in, out, n, nv);
    return 0;
}
