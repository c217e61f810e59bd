// r3_tile.cuh — tiles of TL lanes x L points for L = 3^K (9 <= L <= 2187): a Stockham autosort radix-9/3 FFT with
// E = 9 points per thread in registers and one shared-memory exchange between stages.
//
// rustfft's scalar planner uses its `Radix3` algorithm for lengths 3^k (consumed at
// scirs2-fft/src/fft/algorithms.rs:159-167), so `fft(&x, Some(1_594_323))` (BASELINE configs[3], 3^13) costs
// 5 N log2 N flops there; the padded-convolution route (Bluestein over M = 2^22) spends four times that and
// 16 (2N + 6M) bytes.  With these tiles 3^13 = 729 x 2187 is an ordinary two-pass four-step transform.
//
// Same index algebra as fft_tile.cuh (stage radix R at stride S, S = product of the radices before it):
//   butterfly ib in [0, L/R):  q = ib mod S, base = ib - q
//   reads  x[ib + r*L/R], r < R;   writes y[q + R*base + k*S] = W_L^(base*k) * DFT_R(x)[k]
//   last stage (S*R == L): output index ib + k*L/R  ->  the register pattern of the loads, coalesced stores.
// The I/O description (PassParams / IoDesc), the conjugation flags for inverse transforms, the four-step store twiddle
// and the CTA order are the ones of the power-of-two tile kernel; only full, unmasked tiles are taken (the planner checks).
#pragma once
#include "fft_tile.cuh"
#include "kernel_registry.h"

namespace sfc {

#ifndef SFC_R3_INPLACE_MID
#define SFC_R3_INPLACE_MID 0
#endif

template <typename T>
__device__ __forceinline__ void dft3(Cx<T>& a0, Cx<T>& a1, Cx<T>& a2) {
    constexpr T S3 = (T)0.86602540378443864676372317075294L;  // sin(pi/3)
    const Cx<T> t = cadd(a1, a2);
    const Cx<T> d = csub(a1, a2);
    const Cx<T> m = {fma((T)-0.5, t.x, a0.x), fma((T)-0.5, t.y, a0.y)};
    const Cx<T> s = {S3 * d.y, -(S3 * d.x)};  // -i * sin(pi/3) * d
    a0 = cadd(a0, t);
    a1 = cadd(m, s);
    a2 = csub(m, s);
}

// a * W9^J, W9 = exp(-2*pi*i/9)
template <int J, typename T>
__device__ __forceinline__ Cx<T> mul_w9(Cx<T> a) {
    constexpr T C1 = (T)0.76604444311897803520239265055542L, S1 = (T)0.64278760968653932632264340990726L;   // cos, sin 40 deg
    constexpr T C2 = (T)0.17364817766693034885171662676931L, S2 = (T)0.98480775301220805936674302458952L;   // 80 deg
    constexpr T C4 = (T)-0.93969262078590838405410927732473L, S4 = (T)0.34202014332566873304409961468226L;  // 160 deg
    if constexpr (J == 0) return a;
    else if constexpr (J == 1) return {fma(a.y, S1, a.x * C1), fma(a.y, C1, -(a.x * S1))};
    else if constexpr (J == 2) return {fma(a.y, S2, a.x * C2), fma(a.y, C2, -(a.x * S2))};
    else if constexpr (J == 4) return {fma(a.y, S4, a.x * C4), fma(a.y, C4, -(a.x * S4))};
    else { static_assert(J < 0, "unsupported W9 power"); return a; }
}

// 9-point DFT, natural-order output: n = 3*n1 + n2, k = k1 + 3*k2
template <typename T>
__device__ __forceinline__ void dft9(Cx<T> (&v)[9]) {
#pragma unroll
    for (int n2 = 0; n2 < 3; ++n2) dft3(v[n2], v[n2 + 3], v[n2 + 6]);
    // v[n2 + 3*k1] = y[n2][k1]; twiddle W9^(n2*k1)
    v[1 + 3] = mul_w9<1>(v[1 + 3]);
    v[2 + 3] = mul_w9<2>(v[2 + 3]);
    v[1 + 6] = mul_w9<2>(v[1 + 6]);
    v[2 + 6] = mul_w9<4>(v[2 + 6]);
#pragma unroll
    for (int k1 = 0; k1 < 3; ++k1) dft3(v[3 * k1], v[3 * k1 + 1], v[3 * k1 + 2]);
    // v[k2 + 3*k1] holds X[k1 + 3*k2]
    Cx<T> o[9];
#pragma unroll
    for (int k1 = 0; k1 < 3; ++k1)
#pragma unroll
        for (int k2 = 0; k2 < 3; ++k2) o[k1 + 3 * k2] = v[k2 + 3 * k1];
#pragma unroll
    for (int k = 0; k < 9; ++k) v[k] = o[k];
}

template <typename T, int L_, int TL_>
struct R3Cfg {
    static constexpr int L = L_, TL = TL_, E = 9;
    static constexpr int TPL = L / 9;
    static constexpr int NT = TPL * TL;
    // Exchange layout in slots of one complex element.  A warp-wide 128-bit access is served in groups of eight threads and
    // costs one wavefront per group only when the eight slots differ modulo 8 (ncu source page of the round-2 tiles: every
    // LDS/STS of a strided tile at exactly twice its ideal wavefronts with the plain pitch L).  Every access of the stages
    // is "lane base + i + constant" modulo 8 (9^k = 1 mod 8), so the lane bases decide:
    //   * lane-fastest (column) mapping, tid = TL*i + t: base(t) = r(t) mod 8 with r(t) = (TL mod 8) * t for odd TL
    //     (residue = TL * tid, a bijection on eight consecutive tids) and r(t) = (TL/2 mod 8) * (t >> 1) + 4 * (t & 1) for
    //     even TL (found by exhaustive search over the tiles in use, tools/bank_sim_r3.py);
    //   * butterfly-fastest (row) mapping: base(t) = t * L mod 8, the residues simply continue across a lane boundary.
    // f32 tiles (64-bit accesses, sixteen threads per wavefront) keep the plain pitch.
    static constexpr bool PADDED = sizeof(T) == 8;
    static constexpr int LP = PADDED ? (L + 7) / 8 * 8 : L;
    static __device__ __forceinline__ int lane_base(int t, bool col) {
        if constexpr (!PADDED) return t * LP;
        else {
            const int r = col ? ((TL & 1) ? (TL & 7) * t : ((TL / 2) & 7) * (t >> 1) + 4 * (t & 1)) : t * L;
            return t * LP + (r & 7);
        }
    }
    static constexpr size_t SMEM = ((size_t)TL * LP + (PADDED ? 8 : 0)) * sizeof(Cx<T>);
    // resident CTAs the register allocator leaves room for: three 243-thread tiles (35 KiB) or two 486-thread tiles (70 KiB)
    // per SM — the first cut ran ONE 729-thread CTA per SM (70 registers) at 44 % of its roofline
    static constexpr int MINB = NT <= 256 ? 3 : (NT <= 512 ? 2 : 1);
    static_assert(L % 9 == 0 && NT <= 1024, "L must be a multiple of 9 and the tile at most 1024 threads");
};

template <typename T, typename C, int S>
__device__ __forceinline__ void r3_stages(Cx<T> (&a)[9], Cx<T>* __restrict__ sm, const Cx<T>* __restrict__ tw, int bw, int iw,
                                          int br, int ir) {  // bw / br: lane base (slots) of the writing / reading mapping
    constexpr int L = C::L, TPL = C::TPL;
    constexpr int R = (L / S >= 9) ? 9 : (L / S);  // 9, or a trailing 3
    static_assert(R == 9 || R == 3, "L must be a power of three");
    constexpr bool LAST = (S * R == L);
    constexpr int NB = 9 / R;
    // a middle stage followed by the last one writes IN PLACE (output k into the slot input k came from, touched by this
    // thread only): no barrier against overwriting the previous exchange; the last stage reads through the permutation
    //   Stockham element q + S k + 9 S h  ->  in-place element q + S h + (L / 9) k      (same scheme as fft_tile.cuh)
    // Measured (profiles/r2t_final_single_gpu.log): no gain on these tiles (3^13 x 256 7.20 -> 7.22 ms, 729-point rows 75.8 ->
    // 73.6 %): off, -DSFC_R3_INPLACE_MID=1 builds it.
    constexpr bool MID_INPLACE = SFC_R3_INPLACE_MID && R == 9 && !LAST && S > 1 && L >= 729 && (L / (S * 9) == 3 || L / (S * 9) == 9);  // 243: the permuted reads straddle bank groups (tools/bank_sim_r3.py)
    if constexpr (!LAST && S > 1 && !MID_INPLACE) __syncthreads();  // readers of the previous exchange are done
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        Cx<T> v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = a[b + r * NB];
        if constexpr (R == 9) dft9(v);
        else dft3(v[0], v[1], v[2]);
        if constexpr (LAST) {
#pragma unroll
            for (int k = 0; k < R; ++k) a[b + k * NB] = v[k];
        } else {
            const int ib = iw + b * TPL;
            const int q = ib % S;
            const int base = ib - q;
            // v[k] *= W_L^(base*k): powers of w1 by a short product tree
            const Cx<T> w1 = tw[base];
            const Cx<T> w2 = csqr(w1);
            v[1] = cmul(v[1], w1);
            v[2] = cmul(v[2], w2);
            if constexpr (R == 9) {
                const Cx<T> w3 = cmul(w2, w1), w4 = csqr(w2);
                v[3] = cmul(v[3], w3);
                v[4] = cmul(v[4], w4);
                v[5] = cmul(v[5], cmul(w4, w1));
                v[6] = cmul(v[6], csqr(w3));
                v[7] = cmul(v[7], cmul(w4, w3));
                v[8] = cmul(v[8], csqr(w4));
            }
            if constexpr (MID_INPLACE) {
                Cx<T>* dst = sm + bw + iw;
#pragma unroll
                for (int k = 0; k < R; ++k) dst[k * TPL] = v[k];
            } else {
                Cx<T>* dst = sm + bw + q + R * base;
#pragma unroll
                for (int k = 0; k < R; ++k) dst[k * S] = v[k];
            }
        }
    }
    if constexpr (MID_INPLACE) {
        __syncthreads();
        constexpr int RN = L / (S * 9);  // radix of the last stage
        const Cx<T>* src = sm + br + (ir % S) + TPL * (ir / S);
#pragma unroll
        for (int m = 0; m < 9; ++m) a[m] = src[S * ((RN * m) / 9) + TPL * ((RN * m) % 9)];
        r3_stages<T, C, S * R>(a, sm, tw, br, ir, br, ir);
    } else if constexpr (!LAST) {
        __syncthreads();
        const Cx<T>* src = sm + br + ir;
#pragma unroll
        for (int m = 0; m < 9; ++m) a[m] = src[m * TPL];
        r3_stages<T, C, S * R>(a, sm, tw, br, ir, br, ir);
    }
}

template <typename T, int L, int TL>
__global__ void __launch_bounds__(R3Cfg<T, L, TL>::NT, R3Cfg<T, L, TL>::MINB) r3_tile_kernel(const __grid_constant__ PassParams p) {
    using C = R3Cfg<T, L, TL>;
    using cx = Cx<T>;
    constexpr int E = 9, TPL = C::TPL;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    cx* sm = reinterpret_cast<cx*>(smem_raw);
    const int tid = threadIdx.x;
    uint32_t tile, batch;
    decode_block(p, blockIdx.x, tile, batch);
    int t0, i0, t1, i1;
    map_thread<TL, TPL>(p.map_in, tid, t0, i0);
    map_thread<TL, TPL>(p.map_out, tid, t1, i1);

    cx a[E];
    {
        const uint32_t lane = tile * TL + (uint32_t)t0;
        const bool valid = lane < p.nlanes;  // the last tile of a batch may be partial: its idle lanes only keep the barriers
        const uint32_t lo = lane / p.inner_count, li = lane - lo * p.inner_count;
        const cx* __restrict__ src = reinterpret_cast<const cx*>(p.in.ptr) + (int64_t)batch * p.in.batch_stride +
                                     (int64_t)lo * p.in.outer_stride + (int64_t)li * p.in.inner_stride +
                                     (int64_t)i0 * p.in.elem_stride;
        const int64_t step = (int64_t)TPL * p.in.elem_stride;
#pragma unroll
        for (int m = 0; m < E; ++m) a[m] = valid ? src[m * step] : cx{(T)0, (T)0};
        if (p.flags & F_CONJ_LD_PRE) {
#pragma unroll
            for (int m = 0; m < E; ++m) a[m].y = -a[m].y;
        }
    }
    // column-friendly lane bases when both mappings are lane-fastest, or one of them is and the lanes are long enough that
    // the butterfly-fastest accesses rarely straddle two lanes (tools/bank_sim_r3.py prints every case)
    const bool col = (p.map_in == MAP_COL && p.map_out == MAP_COL) || ((p.map_in == MAP_COL || p.map_out == MAP_COL) && L >= 243);
    const int b0 = C::lane_base(t0, col), b1 = C::lane_base(t1, col);
    r3_stages<T, C, 1>(a, sm, reinterpret_cast<const cx*>(p.tw), b0, i0, b1, i1);
    if constexpr (C::L == 9) {
        if (p.map_in != p.map_out) {  // single-stage tiles never pass through shared memory: remap explicitly
            __syncthreads();
#pragma unroll
            for (int m = 0; m < E; ++m) sm[b0 + m] = a[m];
            __syncthreads();
#pragma unroll
            for (int m = 0; m < E; ++m) a[m] = sm[b1 + m];
        }
    }
    const uint32_t lane = tile * TL + (uint32_t)t1;
    const uint32_t lo = lane / p.inner_count, li = lane - lo * p.inner_count;
    if (p.st_op == ST_TW) fourstep_twiddle<T, E, TPL>(a, p.tw_lo, p.tw_hi, p.tw_shift, (p.flags & F_TW_CONJ) != 0, i1, lo);
    if (p.scale != 1.0) {
        const T sc = (T)p.scale;
#pragma unroll
        for (int m = 0; m < E; ++m) a[m] = {a[m].x * sc, a[m].y * sc};
    }
    if (p.flags & F_CONJ_ST_POST) {
#pragma unroll
        for (int m = 0; m < E; ++m) a[m].y = -a[m].y;
    }
    cx* __restrict__ dst = reinterpret_cast<cx*>(p.out.ptr) + (int64_t)batch * p.out.batch_stride + (int64_t)lo * p.out.outer_stride +
                           (int64_t)li * p.out.inner_stride + (int64_t)i1 * p.out.elem_stride;
    const int64_t ostep = (int64_t)TPL * p.out.elem_stride;
    if (lane < p.nlanes) {
#pragma unroll
        for (int m = 0; m < E; ++m) dst[m * ostep] = a[m];
    }
}

enum : int { TM_R3 = 10 };

template <typename T, int L, int TL>
struct R3Inst {
    using C = R3Cfg<T, L, TL>;
    static cudaError_t launch(const PassParams& p, unsigned grid, cudaStream_t s) {
        static bool configured[64] = {};
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if (dev < 64 && !configured[dev]) {
            e = cudaFuncSetAttribute(r3_tile_kernel<T, L, TL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);
            if (e != cudaSuccess) return e;
            configured[dev] = true;
        }
#ifdef SFC_HOST_EMUL
        emul_launch(&r3_tile_kernel<T, L, TL>, p, grid, C::NT);
#else
        r3_tile_kernel<T, L, TL><<<grid, C::NT, C::SMEM, s>>>(p);
#endif
        return cudaGetLastError();
    }
    static KernelEntry entry() {
        KernelEntry k;
        k.prec = sizeof(T) == 8 ? PREC_F64 : PREC_F32;
        k.L = L;
        k.TL = TL;
        k.E = 9;
        k.dbl = 0;
        k.mode = TM_R3;
        k.groups = 1;
        k.threads = C::NT;
        k.smem = C::SMEM;
        k.func = (const void*)r3_tile_kernel<T, L, TL>;
        k.launch = &launch;
        return k;
    }
};

}  // namespace sfc
