// kernel_inst.cuh — instantiate tile_fft_kernel and wrap it in a launcher.
#pragma once
#include "fft_tile.cuh"
#include "kernel_registry.h"

namespace sfc {

template <typename T, int L, int TL, bool DBL, int EMAX = 16, int MODE = 0, int GROUPS = 1>
struct KernelInst {
    using C = TileCfg<T, L, TL, EMAX, GROUPS, MODE == TM_PIPE_C2C && GROUPS == 1>;
    static constexpr bool GP = MODE == TM_PIPE_C2C && GROUPS == 2;
    static constexpr bool LATE = MODE == TM_PIPE_LATE;
    static constexpr bool PERSISTENT = MODE == TM_PIPE_C2C || LATE;
    static constexpr size_t SMEM_BYTES = GP ? C::SMEM_GP : (LATE ? C::SMEM + 16 : C::SMEM);
    static cudaError_t launch(const PassParams& p0, unsigned grid, cudaStream_t s) {
        PassParams p = p0;
        static bool configured[64] = {};
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 64 && !configured[dev]) {
            e = cudaFuncSetAttribute(tile_fft_kernel<T, L, TL, DBL, EMAX, MODE, GROUPS>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
            if (e != cudaSuccess) return e;
            configured[dev] = true;
        }
        if (PERSISTENT) {
            // persistent: as many CTAs as can be resident (2 per SM), each walking over its share of the tiles
            static int resident[64] = {};
            if (dev < 64 && !resident[dev]) {
                int sms = 0, per_sm = 0;
                cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tile_fft_kernel<T, L, TL, DBL, EMAX, MODE, GROUPS>,
                                                              C::NT, SMEM_BYTES);
                resident[dev] = sms * (per_sm > 0 ? per_sm : 1);
            }
            p.total_tiles = grid;
            const unsigned cap = (unsigned)(dev < 64 ? resident[dev] : 296);
            if (grid > cap) grid = cap;
        }
#ifdef SFC_HOST_EMUL  // tests/emul only: the kernel source run on the host to check index logic without a GPU
        emul_launch(&tile_fft_kernel<T, L, TL, DBL, EMAX, MODE, GROUPS>, p, grid, C::NT);
#else
        tile_fft_kernel<T, L, TL, DBL, EMAX, MODE, GROUPS><<<grid, C::NT, SMEM_BYTES, s>>>(p);
#endif
        return cudaGetLastError();
    }
    static KernelEntry entry() {
        KernelEntry k;
        k.prec = sizeof(T) == 8 ? PREC_F64 : PREC_F32;
        k.L = L;
        k.TL = GP ? C::TLG : TL;  // lanes per SCHEDULED tile: a group-pipelined CTA works on two of them at a time
        k.E = C::E;
        k.dbl = DBL ? 1 : 0;
        k.mode = MODE;
        k.groups = GROUPS;
        k.threads = C::NT;
        k.smem = SMEM_BYTES;
        k.func = (const void*)tile_fft_kernel<T, L, TL, DBL, EMAX, MODE, GROUPS>;
        k.launch = &launch;
        return k;
    }
};

}  // namespace sfc

// generic + fast complex flavour of one tile shape
#define SFC_ADD(T, L, TL, DBL)                              \
    add(::sfc::KernelInst<T, L, TL, DBL, 16, 0>::entry()); \
    add(::sfc::KernelInst<T, L, TL, DBL, 16, 1>::entry());
// fused real-transform flavours (row tiles only)
#define SFC_ADD_REAL(T, L, TL)                                \
    add(::sfc::KernelInst<T, L, TL, false, 16, 2>::entry()); \
    add(::sfc::KernelInst<T, L, TL, false, 16, 3>::entry());
// two independent thread groups per CTA (wide column tiles)
#define SFC_ADD_G2(T, L, TL)                                      \
    add(::sfc::KernelInst<T, L, TL, false, 16, 0, 2>::entry()); \
    add(::sfc::KernelInst<T, L, TL, false, 16, 1, 2>::entry());
// persistent, TMA-pipelined complex flavour (64 KiB tiles: half-size exchange buffer + landing buffer, 2 CTAs / SM)
#define SFC_ADD_PIPE(T, L, TL, DBL) add(::sfc::KernelInst<T, L, TL, DBL, 16, 4>::entry());
// group-pipelined: TL2 = lanes of the two groups together (each group owns TL2 / 2)
#define SFC_ADD_GPIPE(T, L, TL2, DBL) add(::sfc::KernelInst<T, L, TL2, DBL, 16, 4, 2>::entry());
// fused DCT-II rows (Makhoul packing on the half-length transform)
#define SFC_ADD_DCT2(T, L, TL)                               \
    add(::sfc::KernelInst<T, L, TL, false, 16, 5>::entry()); \
    add(::sfc::KernelInst<T, L, TL, false, 16, 6>::entry());
// fused DCT-IV / DST-IV rows (one half-length complex transform, twiddles on load and store)
#define SFC_ADD_DCT4(T, L, TL) add(::sfc::KernelInst<T, L, TL, false, 16, 7>::entry());
// middle pass of the three-pass 2-D plan (16 x L/16 two-dimensional tile)
#define SFC_ADD_2D(T, L, TL) add(::sfc::KernelInst<T, L, TL, false, 16, 8>::entry());
// persistent flavour with the late prefetch into the idle exchange buffer (no extra shared memory)
#define SFC_ADD_PIPE_LATE(T, L, TL) add(::sfc::KernelInst<T, L, TL, false, 16, 9>::entry());
#define SFC_ADD_E(T, L, TL, DBL, E) add(::sfc::KernelInst<T, L, TL, DBL, E>::entry());
