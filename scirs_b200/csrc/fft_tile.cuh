// fft_tile.cuh — the one hot kernel: a tile of TL lanes x L points (L = 2^k)
// transformed by a Stockham autosort radix-16/8/4/2 FFT held in registers
// (E = 16 points per thread) with shared-memory exchanges between stages.
//
// It replaces the arithmetic the reference delegates to rustfft's scalar
// planner (`fft.process(&mut buffer)`, scirs2-fft/src/fft/algorithms.rs:167,
// 362, 376, 683) together with the gather/scatter/scale loops around it
// (:353-395, :677-703).  Forward sign is exp(-2*pi*i*jk/n), unnormalised, as in
// rustfft; the inverse is the same code with the data conjugated on the way in and out.
//
// Index algebra (DIF Stockham, stage radix R at stride S, S*n = L):
//   thread butterfly index ib in [0, L/R):  q = ib mod S, base = ib - q
//   reads  x[ib + r*L/R],            r < R      (always "i + m*L/E": coalesced)
//   writes y[q + R*base + k*S] = W_L^(base*k) * DFT_R(x)[k]
//   last stage (S*R == L): base = 0, output index ib + k*L/R -> same register
//   pattern as the loads, so global stores are coalesced too.
//
// Shared-memory slot of logical element e of lane t:  t*LP + swz(e) with an XOR
// swizzle chosen per exchange so that both the strided writes and the unit-
// stride reads are bank-conflict free (tools/bank_sim.py checks every config).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pass_params.h"

namespace sfc {

template <typename T>
struct alignas(2 * sizeof(T)) Cx {
    T x, y;
};

template <typename T>
__device__ __forceinline__ Cx<T> cadd(Cx<T> a, Cx<T> b) { return {a.x + b.x, a.y + b.y}; }
template <typename T>
__device__ __forceinline__ Cx<T> csub(Cx<T> a, Cx<T> b) { return {a.x - b.x, a.y - b.y}; }
template <typename T>
__device__ __forceinline__ Cx<T> cmul(Cx<T> a, Cx<T> b) {
    return {fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x)};
}
template <typename T>
__device__ __forceinline__ Cx<T> cmulc(Cx<T> a, Cx<T> b) {  // a * conj(b)
    return {fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -(a.x * b.y))};
}
template <typename T>
__device__ __forceinline__ Cx<T> csqr(Cx<T> a) {
    return {fma(a.x, a.x, -(a.y * a.y)), (a.x + a.x) * a.y};
}
template <typename T>
__device__ __forceinline__ Cx<T> cconjf(Cx<T> a) { return {a.x, -a.y}; }
template <typename T>
__device__ __forceinline__ Cx<T> cconj(Cx<T> a) { return {a.x, -a.y}; }
template <typename T>
__device__ __forceinline__ Cx<T> mul_mi(Cx<T> a) { return {a.y, -a.x}; }  // a * (-i)

// ---- global loads with a cache policy ------------------------------------------------------
// Tile data is read exactly once; SFC_LDPOL_DATA / SFC_LDPOL_AUX pick the PTX cache operator
// (0 default, 1 .cg = L2 only, 2 .cs = streaming, 3 L1::no_allocate, 4 .nc read-only path).
#ifndef SFC_LDPOL_DATA
#define SFC_LDPOL_DATA 0
#endif
#ifndef SFC_LDPOL_AUX
#define SFC_LDPOL_AUX 0
#endif
template <int POL>
__device__ __forceinline__ Cx<double> ld_pol(const Cx<double>* p) {
    Cx<double> r;
    if constexpr (POL == 1) asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    else if constexpr (POL == 2) asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    else if constexpr (POL == 3) asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    else if constexpr (POL == 4) asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    else r = *p;
    return r;
}
template <int POL>
__device__ __forceinline__ Cx<float> ld_pol(const Cx<float>* p) {
    Cx<float> r;
    if constexpr (POL == 1) asm volatile("ld.global.cg.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    else if constexpr (POL == 2) asm volatile("ld.global.cs.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    else if constexpr (POL == 3) asm volatile("ld.global.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    else if constexpr (POL == 4) asm volatile("ld.global.nc.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    else r = *p;
    return r;
}

// ---- f32: Blackwell packed f32x2 arithmetic -----------------------------------------------
// sm_100 has FADD2 / FMUL2 / FFMA2 on 64-bit register pairs with free half-swap / broadcast operand
// modifiers, so one complex f32 add is ONE instruction and the -i rotation folds into the consumer.
// These overloads are picked for Cx<float>; the generic templates above stay for f64.
typedef unsigned long long u64p;
__device__ __forceinline__ u64p pk2(float lo, float hi) {
    u64p r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ u64p pk2(Cx<float> a) { return pk2(a.x, a.y); }
__device__ __forceinline__ Cx<float> upk2(u64p v) {
    Cx<float> r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ u64p add2(u64p a, u64p b) {
    u64p r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64p sub2(u64p a, u64p b) {
    u64p r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64p mul2(u64p a, u64p b) {
    u64p r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64p fma2(u64p a, u64p b, u64p c) {
    u64p r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ Cx<float> cadd(Cx<float> a, Cx<float> b) { return upk2(add2(pk2(a), pk2(b))); }
__device__ __forceinline__ Cx<float> csub(Cx<float> a, Cx<float> b) { return upk2(sub2(pk2(a), pk2(b))); }
// The half-negated swaps below ((-t.y, t.x) and (u.y, -u.x)) are written so that they feed an ADD:
// ptxas folds them into FADD2's operand modifier (R.F32x2.LO_HI.NP); FMUL2 / FFMA2 cannot take them.
__device__ __forceinline__ Cx<float> cmul(Cx<float> a, Cx<float> b) {
    // ax*b + i*(ay*b)
    const Cx<float> t = upk2(mul2(pk2(a.y, a.y), pk2(b)));
    return upk2(add2(mul2(pk2(a.x, a.x), pk2(b)), pk2(-t.y, t.x)));
}
__device__ __forceinline__ Cx<float> cmulc(Cx<float> a, Cx<float> b) {  // a * conj(b)
    // bx*a - i*(by*a)
    const Cx<float> u = upk2(mul2(pk2(b.y, b.y), pk2(a)));
    return upk2(add2(mul2(pk2(b.x, b.x), pk2(a)), pk2(u.y, -u.x)));
}
__device__ __forceinline__ Cx<float> csqr(Cx<float> a) { return cmul(a, a); }

// ---- radix butterflies (forward, in place, natural-order output) ------------

template <typename T>
__device__ __forceinline__ void dft2(Cx<T>& a0, Cx<T>& a1) {
    Cx<T> t = a0;
    a0 = cadd(t, a1);
    a1 = csub(t, a1);
}

template <typename T>
__device__ __forceinline__ void dft4(Cx<T>& a0, Cx<T>& a1, Cx<T>& a2, Cx<T>& a3) {
    Cx<T> t0 = cadd(a0, a2), t1 = csub(a0, a2);
    Cx<T> t2 = cadd(a1, a3), t3 = mul_mi(csub(a1, a3));
    a0 = cadd(t0, t2);
    a1 = cadd(t1, t3);
    a2 = csub(t0, t2);
    a3 = csub(t1, t3);
}

// a * W16^J  (W16 = exp(-2*pi*i/16)), J folded at compile time
template <int J, typename T>
__device__ __forceinline__ Cx<T> mul_w16(Cx<T> a) {
    constexpr T H = (T)0.70710678118654752440084436210485L;
    constexpr T C1 = (T)0.92387953251128675612818318939679L;  // cos(pi/8)
    constexpr T S1 = (T)0.38268343236508977172845998403040L;  // sin(pi/8)
    if constexpr (J == 0) return a;
    else if constexpr (J == 1) return {fma(a.y, S1, a.x * C1), fma(a.y, C1, -(a.x * S1))};
    else if constexpr (J == 2) return {(a.x + a.y) * H, (a.y - a.x) * H};
    else if constexpr (J == 3) return {fma(a.y, C1, a.x * S1), fma(a.y, S1, -(a.x * C1))};
    else if constexpr (J == 4) return {a.y, -a.x};
    else if constexpr (J == 6) return {(a.y - a.x) * H, -((a.x + a.y) * H)};
    else if constexpr (J == 9) return {-fma(a.y, S1, a.x * C1), fma(a.x, S1, -(a.y * C1))};
    else { static_assert(J < 0, "unsupported W16 power"); return a; }
}

// a * W16^J for f32 with packed operations
template <int J>
__device__ __forceinline__ Cx<float> mul_w16(Cx<float> a) {
    constexpr float H = 0.70710678118654752440f, C1 = 0.92387953251128675613f, S1 = 0.38268343236508977173f;
    if constexpr (J == 0) return a;
    else if constexpr (J == 4) return {a.y, -a.x};
    else if constexpr (J == 2) return upk2(mul2(add2(pk2(a), pk2(a.y, -a.x)), pk2(H, H)));
    else if constexpr (J == 6) return upk2(mul2(sub2(pk2(a.y, -a.x), pk2(a)), pk2(H, H)));
    else {
        // a * (c - i s) = c*a - i*(s*a)
        constexpr float c = J == 1 ? C1 : (J == 3 ? S1 : -C1);
        constexpr float sn = J == 1 ? S1 : (J == 3 ? C1 : -S1);
        static_assert(J == 1 || J == 3 || J == 9, "unsupported W16 power");
        const Cx<float> u = upk2(mul2(pk2(sn, sn), pk2(a)));
        return upk2(add2(mul2(pk2(c, c), pk2(a)), pk2(u.y, -u.x)));
    }
}

template <int R, typename T>
struct Dft;

template <typename T>
struct Dft<2, T> {
    static __device__ __forceinline__ void run(Cx<T> (&v)[2]) { dft2(v[0], v[1]); }
};
template <typename T>
struct Dft<4, T> {
    static __device__ __forceinline__ void run(Cx<T> (&v)[4]) { dft4(v[0], v[1], v[2], v[3]); }
};
template <typename T>
struct Dft<8, T> {
    // n = 4*n1 + n2, k = k1 + 2*k2
    static __device__ __forceinline__ void run(Cx<T> (&v)[8]) {
        Cx<T> y0[4], y1[4];
#pragma unroll
        for (int n2 = 0; n2 < 4; ++n2) {
            y0[n2] = cadd(v[n2], v[n2 + 4]);
            y1[n2] = csub(v[n2], v[n2 + 4]);
        }
        y1[1] = mul_w16<2>(y1[1]);
        y1[2] = mul_w16<4>(y1[2]);
        y1[3] = mul_w16<6>(y1[3]);
        dft4(y0[0], y0[1], y0[2], y0[3]);
        dft4(y1[0], y1[1], y1[2], y1[3]);
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {
            v[2 * k2] = y0[k2];
            v[2 * k2 + 1] = y1[k2];
        }
    }
};
template <typename T>
struct Dft<16, T> {
    // n = 4*n1 + n2, k = k1 + 4*k2
    static __device__ __forceinline__ void run(Cx<T> (&v)[16]) {
#pragma unroll
        for (int n2 = 0; n2 < 4; ++n2) dft4(v[n2], v[n2 + 4], v[n2 + 8], v[n2 + 12]);
        // v[n2 + 4*k1] now holds y[n2][k1]; apply W16^(n2*k1)
        v[1 + 4] = mul_w16<1>(v[1 + 4]);
        v[2 + 4] = mul_w16<2>(v[2 + 4]);
        v[3 + 4] = mul_w16<3>(v[3 + 4]);
        v[1 + 8] = mul_w16<2>(v[1 + 8]);
        v[2 + 8] = mul_w16<4>(v[2 + 8]);
        v[3 + 8] = mul_w16<6>(v[3 + 8]);
        v[1 + 12] = mul_w16<3>(v[1 + 12]);
        v[2 + 12] = mul_w16<6>(v[2 + 12]);
        v[3 + 12] = mul_w16<9>(v[3 + 12]);
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
        // v[k2 + 4*k1] holds X[k1 + 4*k2]: transpose register names
        Cx<T> o[16];
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2) o[k1 + 4 * k2] = v[k2 + 4 * k1];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = o[k];
    }
};

// v[k] *= w^k, k = 1..R-1, powers built by a depth <= 5 product tree
template <int R, typename T>
__device__ __forceinline__ void apply_twiddle_powers(Cx<T> (&v)[R], Cx<T> w1) {
    if constexpr (R >= 2) v[1] = cmul(v[1], w1);
    if constexpr (R >= 4) {
        Cx<T> w2 = csqr(w1);
        v[2] = cmul(v[2], w2);
        Cx<T> w3 = cmul(w2, w1);
        v[3] = cmul(v[3], w3);
        if constexpr (R >= 8) {
            Cx<T> w4 = csqr(w2);
            v[4] = cmul(v[4], w4);
            v[5] = cmul(v[5], cmul(w4, w1));
            v[6] = cmul(v[6], cmul(w4, w2));
            v[7] = cmul(v[7], cmul(w4, w3));
            if constexpr (R >= 16) {
                Cx<T> w8 = csqr(w4);
                v[8] = cmul(v[8], w8);
                v[9] = cmul(v[9], cmul(w8, w1));
                v[10] = cmul(v[10], cmul(w8, w2));
                v[11] = cmul(v[11], cmul(w8, w3));
                Cx<T> w12 = cmul(w8, w4);
                v[12] = cmul(v[12], w12);
                v[13] = cmul(v[13], cmul(w12, w1));
                v[14] = cmul(v[14], cmul(w12, w2));
                v[15] = cmul(v[15], cmul(w12, w3));
            }
        }
    }
}

// v[k] *= W_L^(base*k) read straight from the stage table, for stages whose stride S is >= 16: all threads of a
// half warp then share `base`, every load touches one or two addresses (L1 broadcast), and the 53-instruction FP64
// product tree above is saved.  Measured (tools/r1_experiments/exp31.sh, 65536 x 4096): 1.5% SLOWER for f64 c2c (91.5% against 93.0%
// of the HBM peak) and 3% slower for f64 rfft -- the 15 extra L1 loads per butterfly cost more issue slots next to the
// tile's own global loads than the FP64 tree does.  Off; -DSFC_TW_LOAD=1 builds it.
#ifndef SFC_INPLACE_MID
#define SFC_INPLACE_MID 1
#endif
#ifndef SFC_INPLACE_MID_HOOK  // the in-place middle stage in the late-prefetch (TM_PIPE_LATE) flavour too
#define SFC_INPLACE_MID_HOOK 1
#endif
#ifndef SFC_TW_LOAD
#define SFC_TW_LOAD 0
#endif
template <int R, typename T>
__device__ __forceinline__ void apply_twiddle_table(Cx<T> (&v)[R], const Cx<T>* __restrict__ tw, int base) {
#pragma unroll
    for (int k = 1; k < R; ++k) v[k] = cmul(v[k], tw[base * k]);
}

__host__ __device__ constexpr int ilog2(int x) { return x <= 1 ? 0 : 1 + ilog2(x >> 1); }

// ---- tile configuration ------------------------------------------------------

// GROUPS > 1 splits the CTA into independent thread groups, each owning TL/GROUPS lanes and its own
// named barrier: the groups drift apart like separate CTAs (load of one overlaps compute of the
// other) while the tile keeps TL adjacent lanes, i.e. full 128 B rows, in one CTA at one time.
// SPLIT_ exchanges the real and the imaginary parts in two rounds through a buffer of HALF the size (elements
// of sizeof(T) instead of sizeof(Cx<T>)): the room this frees holds the TMA landing buffer of the pipelined
// flavour (TM_PIPE_C2C), which prefetches the next tile while this one is transformed.
template <typename T, int L_, int TL_, int EMAX_ = 16, int GROUPS_ = 1, bool SPLIT_ = false>
struct TileCfg {
    static constexpr int L = L_;
    static constexpr int TL = TL_;
    static constexpr int GROUPS = GROUPS_;
    static constexpr bool SPLIT = SPLIT_;
    static constexpr int XBYTES = SPLIT_ ? (int)sizeof(T) : (int)sizeof(Cx<T>);  // bytes per exchange slot
    static constexpr int TLG = TL_ / GROUPS_;                 // lanes per group
    static constexpr int E = L < EMAX_ ? L : EMAX_;    // points per thread (= largest radix)
    static constexpr int TPL = L / E;                  // threads per lane
    static constexpr int NT = TPL * TL;                // threads per CTA
    static constexpr int G = 128 / XBYTES;  // threads per smem wavefront
    static __host__ __device__ constexpr int lane_pitch() {
        int want = (TLG < G) ? (G / TLG) % G : 1;
        // padded exchange layout (one pad slot per E elements) and >= L+1 because the
        // real-transform paths stage L+1 points per lane
        int lp = L + (L > E ? L / (E >= 16 ? 8 : E) : 0) + 1;  // E=16 plans may start with a radix-8 stage
        while (lp % G != want) ++lp;
        return lp;
    }
    static constexpr int LP = lane_pitch();
    static constexpr size_t XCH_BYTES = (((size_t)TL * LP * XBYTES) + 127) / 128 * 128;
    static constexpr size_t LAND_BYTES = (size_t)TL * L * sizeof(Cx<T>);  // dense tile, TMA destination
    // pipelined flavour: exchange buffer | landing buffer | one mbarrier
    static constexpr size_t SMEM = SPLIT_ ? XCH_BYTES + LAND_BYTES + 16 : (size_t)TL * LP * sizeof(Cx<T>);
    // group-pipelined flavour (TM_PIPE_C2C with GROUPS == 2): full-size exchange buffers for both groups, ONE landing
    // buffer of a group's tile (TLG lanes) that the groups consume alternately, two mbarriers
    static constexpr size_t LAND_G_BYTES = (size_t)(TL_ / GROUPS_) * L * sizeof(Cx<T>);
    static constexpr size_t SMEM_GP = XCH_BYTES + LAND_G_BYTES + 32;
    // resident CTAs per SM the register allocator must leave room for
    static constexpr int NTG = NT / GROUPS;                   // threads per group
    static __device__ __forceinline__ void sync(int group) {
        if constexpr (GROUPS == 1) {
            __syncthreads();
        } else {
#ifdef SFC_HOST_EMUL
            emul_unsupported("named barriers (two thread groups per CTA)");
#else
            asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(NTG) : "memory");
#endif
        }
    }
#ifndef SFC_MINB_F64
#define SFC_MINB_F64 2  // 3 (85 registers) spills ~0.5 KB per thread: measured slower
#endif
    static constexpr int MINB =
        (E <= 8) ? (NT <= 512 ? (sizeof(T) == 8 ? 2 : 3) : 1)
                 : ((NT <= 128 && sizeof(T) == 8) ? 2 * SFC_MINB_F64
                                                  : ((NT <= 256 && sizeof(T) == 8) ? SFC_MINB_F64 : (NT <= 128 ? 6 : (NT <= 256 ? 3 : 1))));
};

// Padded exchange layout: logical element e of a lane lives at slot e + (e >> log2(R*S)) * PADW,
// PADW = S when S is smaller than a shared-memory wavefront (G slots), else 0.  All offsets a
// thread needs are then "thread base + compile-time constant", so LDS/STS carry immediates.
template <typename C, int R, int S>
struct Xch {
    static constexpr int BLK = R * S;
    static constexpr int SH = ilog2(BLK);
    static constexpr int PADW = (S < C::G) ? S : 0;
    // slot of the first output (k = 0) of butterfly ib; output k sits k*S slots further
    static __device__ __forceinline__ int write_base(int ib) {
        const int q = ib & (S - 1);
        const int p0 = q + R * (ib - q);
        return p0 + (p0 >> SH) * PADW;
    }
    // thread base of the read pattern e = i + m*TPL
    static __device__ __forceinline__ int read_base(int i) {
        if constexpr (C::TPL >= BLK) return i + (i >> SH) * PADW;
        else return i;
    }
    static __host__ __device__ constexpr int read_off(int m) {
        return (C::TPL >= BLK) ? m * (C::TPL + (C::TPL >> SH) * PADW) : m * C::TPL + ((m * C::TPL) >> SH) * PADW;
    }
};

// One Stockham stage of radix R at stride S over the thread's E registers, then
// recurse.  (tw, iw) = mapping of the thread while it holds the stage inputs,
// (tr, ir) = mapping used after the exchange.
// MIRROR (real-to-complex transforms whose last stage has two radix-E/2 butterflies per thread):
// the thread takes butterfly i and its mirror image S_last - i, so that after the last stage the
// pairs (k, L-k) the Hermitian post-pass combines are both in its own registers.
// MIRROR_IN (complex-to-real transforms, same lengths): the radix-8 stage comes FIRST and the thread
// again owns butterfly i and its mirror, so the Hermitian pre-pass pairs X[k], X[L-k] come straight
// from the thread's own global loads.
// what the late-prefetch flavour does once the exchange buffer is idle (after the last exchange has been read)
struct LateIssue {
    const PassParams* p;
    uint32_t next_blk;
    uint32_t bar;
    uint32_t land;
};
template <typename T, int L, int TL>
__device__ __forceinline__ void late_issue(const LateIssue& h);

template <typename T, typename C, int S, bool FIRST, bool MIRROR = false, bool MIRROR_IN = false, bool NOTW1 = false, bool HOOK = false>
__device__ __forceinline__ void run_stages(Cx<T> (&a)[C::E], Cx<T>* __restrict__ sm,
                                           const Cx<T>* __restrict__ tw, int tw_, int iw, int tr,
                                           int ir, int grp = 0, const LateIssue* hook = nullptr) {
    constexpr int L = C::L, E = C::E, TPL = C::TPL;
    constexpr int REM = 1 << (ilog2(L) % ilog2(E));  // the one radix smaller than E (1 = none)
    constexpr int R = MIRROR_IN ? ((S == 1 && REM > 1) ? REM : E) : ((L / S >= E) ? E : (L / S));
    constexpr bool LAST = (S * R == L);
    constexpr int NB = E / R;  // butterflies per thread in this stage
    // Middle stage of the three-stage tiles (L = R1 * 16 * R3; R1 = 16, or 8 in the complex-to-real flavour) done IN PLACE: output
    // k goes back to the slot input k came from, which only this thread reads or writes, so the barrier that protects the
    // exchange buffer against overwriting (one of the three barriers of a tile) is not needed; the last stage then reads through
    // the permuted positions:  Stockham element q + S k + 16 S h  ->  in-place element q + S h + (L / 16) k   (S = R1).
    // The streaming proxy of the tile (tools/micro/power_roofline.cu) puts ~3 points of the HBM roofline on every CTA-wide
    // barrier; measured on the tiles (profiles/r2r_inplace_mid.log): 4096-point c2c 78.7 -> 81.2 %, 2048 84.4 -> 86.1 %.
    // Real flavours (65,536 x 4096, sustained, profiles/r2s_inplace_real.log): complex-to-real f64 85.9 -> 87.4 % (kept), f32
    // 79.1 -> 77.7 % (off); real-to-complex with the mirrored last stage f64 84.9 -> 83.6 % (a 12-byte spill; off), f32 +0.5.
    constexpr bool MID_INPLACE = SFC_INPLACE_MID && (!HOOK || SFC_INPLACE_MID_HOOK) && !C::SPLIT && E == 16 && R == 16 && !LAST && S > 1 &&
                                 (SFC_INPLACE_MID > 1 || (!MIRROR && (!MIRROR_IN || sizeof(T) == 8))) &&
                                 (MIRROR_IN ? S == 8 : S == 16) && L / (S * R) >= 2 && L / (S * R) <= 16;
    if constexpr (!LAST && !FIRST && !MID_INPLACE) C::sync(grp);  // previous readers done before we overwrite
    if constexpr (C::SPLIT && !LAST) {
        // half-size exchange buffer: real parts first, imaginary parts second (outputs parked in a[] meanwhile)
        static_assert(!MIRROR && !MIRROR_IN, "split exchange is for the plain complex flavours");
        T* __restrict__ smt = reinterpret_cast<T*>(sm);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            Cx<T> v[R];
#pragma unroll
            for (int r = 0; r < R; ++r) v[r] = a[b + r * NB];
            Dft<R, T>::run(v);
            if constexpr (NOTW1 && S == 1) {
            } else if constexpr (SFC_TW_LOAD && S >= 16) apply_twiddle_table<R, T>(v, tw, (iw + b * TPL) & ~(S - 1));
            else apply_twiddle_powers<R, T>(v, tw[(iw + b * TPL) & ~(S - 1)]);
#pragma unroll
            for (int k = 0; k < R; ++k) a[b + k * NB] = v[k];
        }
        const T* src = smt + tr * C::LP + Xch<C, R, S>::read_base(ir);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            T* dst = smt + tw_ * C::LP + Xch<C, R, S>::write_base(iw + b * TPL);
#pragma unroll
            for (int k = 0; k < R; ++k) dst[k * S] = a[b + k * NB].x;
        }
        C::sync(grp);
#pragma unroll
        for (int m = 0; m < E; ++m) a[m].x = src[Xch<C, R, S>::read_off(m)];
        C::sync(grp);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            T* dst = smt + tw_ * C::LP + Xch<C, R, S>::write_base(iw + b * TPL);
#pragma unroll
            for (int k = 0; k < R; ++k) dst[k * S] = a[b + k * NB].y;
        }
        C::sync(grp);
#pragma unroll
        for (int m = 0; m < E; ++m) a[m].y = src[Xch<C, R, S>::read_off(m)];
        run_stages<T, C, S * R, false, MIRROR, MIRROR_IN>(a, sm, tw, tr, ir, tr, ir, grp);
        return;
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        Cx<T> v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = a[b + r * NB];
        Dft<R, T>::run(v);
        if constexpr (LAST) {
#pragma unroll
            for (int k = 0; k < R; ++k) a[b + k * NB] = v[k];
        } else {
            int ib = iw + b * TPL;
            if constexpr (MIRROR_IN && S == 1 && NB == 2) {
                if (b == 1) ib = iw == 0 ? (L / R) / 2 : (L / R) - iw;  // the mirror butterfly
            }
            if constexpr (NOTW1 && S == 1) {
            } else if constexpr (SFC_TW_LOAD && S >= 16) apply_twiddle_table<R, T>(v, tw, ib & ~(S - 1));
            else apply_twiddle_powers<R, T>(v, tw[ib & ~(S - 1)]);
            if constexpr (MID_INPLACE) {
                using X1 = Xch<C, S, 1>;  // the layout the first stage wrote and this thread just read through
                Cx<T>* dst = sm + tw_ * C::LP + X1::read_base(iw);
#pragma unroll
                for (int k = 0; k < R; ++k) dst[X1::read_off(k)] = v[k];
            } else {
                Cx<T>* dst = sm + tw_ * C::LP + Xch<C, R, S>::write_base(ib);
#pragma unroll
                for (int k = 0; k < R; ++k) dst[k * S] = v[k];
            }
        }
    }
    if constexpr (MID_INPLACE) {
        C::sync(grp);
        constexpr int RN = L / (S * R);        // radix of the last stage
        constexpr int KP = TPL + TPL / S;      // slot distance of consecutive k (one pad slot per S elements)
        constexpr int LS = ilog2(S);
        const Cx<T>* lane = sm + tr * C::LP;
        // slot of Stockham element ib + 16 S r (ib < 16 S): (ib mod S) + KP (ib div S) + (S + 1) r
        const int b0 = (ir & (S - 1)) + KP * (ir >> LS);
        if constexpr (MIRROR && E / RN == 2) {
            // butterfly 0 = ir, butterfly 1 = its mirror (16 S - ir, or 8 S for ir = 0)
            const int j2 = ir == 0 ? (S * R) / 2 : S * R - ir;
            const int b1 = (j2 & (S - 1)) + KP * (j2 >> LS);
#pragma unroll
            for (int r = 0; r < RN; ++r) {
                a[2 * r] = lane[b0 + (S + 1) * r];
                a[2 * r + 1] = lane[b1 + (S + 1) * r];
            }
        } else {
#pragma unroll
            for (int m = 0; m < E; ++m) a[m] = lane[b0 + (S + 1) * ((RN * m) >> 4) + KP * ((RN * m) & 15)];
        }
        if constexpr (HOOK) {
            // the last stage never touches shared memory again: the buffer is free for the next tile's data (late prefetch)
            C::sync(grp);
            late_issue<T, C::L, C::TL>(*hook);
        }
        run_stages<T, C, S * R, false, MIRROR, MIRROR_IN, false, HOOK>(a, sm, tw, tr, ir, tr, ir, grp, hook);
    } else if constexpr (!LAST) {
        C::sync(grp);
        const Cx<T>* src = sm + tr * C::LP + Xch<C, R, S>::read_base(ir);
        constexpr int SN = S * R;                                  // stride of the next stage
        constexpr int RN = MIRROR_IN ? E : ((L / SN >= E) ? E : (L / SN));  // its radix
        if constexpr (MIRROR && SN * RN == L && E / RN == 2) {
            // butterfly 0 = ir (elements ir + r*SN), butterfly 1 = its mirror (SN - ir, or SN/2 for ir = 0)
            const Cx<T>* lane = sm + tr * C::LP;
            const int j2 = ir == 0 ? SN / 2 : SN - ir;
#pragma unroll
            for (int r = 0; r < RN; ++r) {
                a[2 * r] = src[Xch<C, R, S>::read_off(2 * r)];
                const int e = j2 + r * SN;
                a[2 * r + 1] = lane[e + (e >> Xch<C, R, S>::SH) * Xch<C, R, S>::PADW];
            }
        } else {
#pragma unroll
            for (int m = 0; m < E; ++m) a[m] = src[Xch<C, R, S>::read_off(m)];
        }
        if constexpr (HOOK && SN * RN == L) {
            // the stage that follows is the last one and never touches shared memory again: once every thread holds its
            // elements the buffer is free for the next tile's data
            C::sync(grp);
            late_issue<T, C::L, C::TL>(*hook);
        }
        run_stages<T, C, S * R, false, MIRROR, MIRROR_IN, false, HOOK>(a, sm, tw, tr, ir, tr, ir, grp, hook);
    }
}

// W_32^m = exp(-2*pi*i*m/32), m < 16 (real-transform post/pre twiddle split)
template <typename T>
__device__ __forceinline__ Cx<T> w32(int m) {
    switch (m) {
        case 0: return {(T)1.0L, (T)-0.0L};
        case 1: return {(T)0.98078528040323044912618223613424L, (T)-0.19509032201612826784828486847702L};
        case 2: return {(T)0.92387953251128675612818318939679L, (T)-0.38268343236508977172845998403040L};
        case 3: return {(T)0.83146961230254523707878837761791L, (T)-0.55557023301960222474283081394853L};
        case 4: return {(T)0.70710678118654752440084436210485L, (T)-0.70710678118654752440084436210485L};
        case 5: return {(T)0.55557023301960222474283081394853L, (T)-0.83146961230254523707878837761791L};
        case 6: return {(T)0.38268343236508977172845998403040L, (T)-0.92387953251128675612818318939679L};
        case 7: return {(T)0.19509032201612826784828486847702L, (T)-0.98078528040323044912618223613424L};
        case 8: return {(T)0.0L, (T)-1.0L};
        case 9: return {(T)-0.19509032201612826784828486847702L, (T)-0.98078528040323044912618223613424L};
        case 10: return {(T)-0.38268343236508977172845998403040L, (T)-0.92387953251128675612818318939679L};
        case 11: return {(T)-0.55557023301960222474283081394853L, (T)-0.83146961230254523707878837761791L};
        case 12: return {(T)-0.70710678118654752440084436210485L, (T)-0.70710678118654752440084436210485L};
        case 13: return {(T)-0.83146961230254523707878837761791L, (T)-0.55557023301960222474283081394853L};
        case 14: return {(T)-0.92387953251128675612818318939679L, (T)-0.38268343236508977172845998403040L};
        default: return {(T)-0.98078528040323044912618223613424L, (T)-0.19509032201612826784828486847702L};
    }
}

// Bluestein chirp values chirp[P + m*D], m < E, chirp[n] = exp(-i*pi*n^2/N), without a table of N
// entries: the phase n^2 mod 2N is reduced exactly in integers and looked up in a two-level root
// table (2 * ~sqrt(2N) entries, cache resident); consecutive positions follow the second-order
// recurrence c[m+1] = c[m] * r[m], r[m+1] = r[m] * q with r[0] = R(2PD + D^2), q = R(2D^2).  Two
// anchored half-chains keep the rounding error at the table's own level (~1e-16).
// x mod m for 0 <= x < 2^52, m < 2^32, in double arithmetic (exact: one fused multiply-add and two fix-ups
// replace the ~150-instruction 64-bit integer remainder)
__device__ __forceinline__ uint32_t mod_small(uint64_t x, double m, double inv_m) {
    const double xd = (double)x;
    const double q = floor(xd * inv_m);
    double r = fma(-q, m, xd);
    if (r < 0.0) r += m;
    if (r >= m) r -= m;
    return (uint32_t)r;
}

template <int E>
__device__ __forceinline__ void chirp_gen(Cx<double> (&c)[E], const PassParams& p, uint64_t P, uint64_t D,
                                          Cx<double> q) {
    const Cx<double>* __restrict__ lo = reinterpret_cast<const Cx<double>*>(p.chirp_lo);
    const Cx<double>* __restrict__ hi = reinterpret_cast<const Cx<double>*>(p.chirp_hi);
    const uint32_t mask = (1u << p.chirp_shift) - 1u;
    auto root = [&](uint32_t k) { return cmul(hi[k >> p.chirp_shift], lo[k & mask]); };
    const double m = (double)p.chirp_mod, inv_m = 1.0 / m;
    constexpr int H = E >= 8 ? E / 2 : E;
    // all products below stay under 2^52 (positions < 2^26: the planner checks)
    const uint64_t dd = D * D;
    c[0] = root(mod_small(P * P, m, inv_m));
    Cx<double> r = root(mod_small(2 * P * D + dd, m, inv_m));
#pragma unroll
    for (int k = 1; k < H; ++k) {
        c[k] = cmul(c[k - 1], r);
        r = cmul(r, q);
    }
    if constexpr (E >= 8) {
        const uint64_t P2 = P + (uint64_t)H * D;
        c[H] = root(mod_small(P2 * P2, m, inv_m));
        r = root(mod_small(2 * P2 * D + dd, m, inv_m));
#pragma unroll
        for (int k = H + 1; k < E; ++k) {
            c[k] = cmul(c[k - 1], r);
            r = cmul(r, q);
        }
    }
}

template <typename T, int E>
__device__ __forceinline__ void chirp_apply(Cx<T> (&a)[E], const PassParams& p, int64_t P, int64_t D, const double* q) {
    Cx<double> c[E];
    chirp_gen<E>(c, p, (uint64_t)P, (uint64_t)D, Cx<double>{q[0], q[1]});
#pragma unroll
    for (int m = 0; m < E; ++m) a[m] = cmul(a[m], Cx<T>{(T)c[m].x, (T)c[m].y});
}

// four-step twiddle W_M^(e*lo), e = i + m*TPL:  w_m = W^(i*lo) * (W^(TPL*lo))^m.  Two two-level table lookups per
// thread, then a depth-4 product tree (no per-element loads).
template <typename T, int E, int TPL>
__device__ __forceinline__ void fourstep_twiddle(Cx<T> (&a)[E], const void* lo_tab, const void* hi_tab, int shift,
                                                 bool conj, int i, uint32_t lo) {
    using cx = Cx<T>;
    const cx* __restrict__ tlo = reinterpret_cast<const cx*>(lo_tab);
    const cx* __restrict__ thi = reinterpret_cast<const cx*>(hi_tab);
    const uint64_t lmask = ((uint64_t)1 << shift) - 1;
    const T sgn = conj ? (T)-1 : (T)1;
    const uint64_t e0 = (uint64_t)i * (uint64_t)lo, e1 = (uint64_t)TPL * (uint64_t)lo;
    cx b = cmul(thi[e0 >> shift], tlo[e0 & lmask]);
    cx s1 = cmul(thi[e1 >> shift], tlo[e1 & lmask]);
    b.y *= sgn;
    s1.y *= sgn;
    if constexpr (E == 16) {
        const cx s2 = csqr(s1), s4 = csqr(s2), s8 = csqr(s4);
        cx w[8];
        w[0] = b;
        w[1] = cmul(b, s1);
        w[2] = cmul(b, s2);
        w[3] = cmul(w[1], s2);
#pragma unroll
        for (int j = 0; j < 4; ++j) w[4 + j] = cmul(w[j], s4);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            a[j] = cmul(a[j], w[j]);
            a[8 + j] = cmul(a[8 + j], cmul(w[j], s8));
        }
    } else {
        cx w = b;
#pragma unroll
        for (int m = 0; m < E; ++m) {
            a[m] = cmul(a[m], w);
            w = cmul(w, s1);
        }
    }
}

template <int TL, int TPL>
__device__ __forceinline__ void map_thread(int mode, int tid, int& t, int& i) {
    if (mode == MAP_COL) {
        t = tid % TL;
        i = tid / TL;
    } else {
        i = tid % TPL;
        t = tid / TPL;
    }
}

// ---- the kernel ---------------------------------------------------------------

#ifdef SFC_PHASE_TIMING
// developer build: thread 0 of every CTA adds the clocks spent between phase marks to p.dbg[k]
#define SFC_PHASE(k)                                                                  \
    do {                                                                              \
        if (p.dbg && threadIdx.x == 0) {                                              \
            const long long now_ = clock64();                                         \
            atomicAdd(p.dbg + (k), (unsigned long long)(now_ - phase_t_));            \
            phase_t_ = now_;                                                          \
        }                                                                             \
    } while (0)
#define SFC_PHASE_FORCE(arr)                                                          \
    do {                                                                              \
        if (p.dbg && threadIdx.x == 0) {                                              \
            double s_ = 0;                                                            \
            _Pragma("unroll") for (int m_ = 0; m_ < E; ++m_) s_ += (double)arr[m_].x; \
            if (s_ == 1.2345e300) atomicAdd(p.dbg + 15, 1ull);                        \
        }                                                                             \
    } while (0)
#else
#define SFC_PHASE(k) do { } while (0)
#define SFC_PHASE_FORCE(arr) do { } while (0)
#endif

// Kernel flavours.  The planner picks a FAST flavour whenever every lane of every tile is valid
// and no bounds mask is needed (PlanBuilder::finish_tile); they carry no predicates, no zero
// fill and no runtime-selected code on the hot path.
enum TileMode : int {
    TM_GENERIC = 0,   // every load / store operator, masks, partial tiles
    TM_FAST_C2C = 1,  // complex in, complex out, optional table multiplies / twiddles / scale
    TM_FAST_R2C = 2,  // packed real rows in, N/2+1 Hermitian half out (fused post-twiddle)
    TM_FAST_C2R = 3,  // N/2+1 Hermitian half in (fused pre-twiddle), packed real rows out
    // TM_FAST_C2C made persistent and software-pipelined: one CTA walks over tiles blockIdx.x, +gridDim.x, ...;
    // while it transforms a tile the TMA unit (cp.async.bulk + mbarrier) lands the next one in shared memory,
    // so the global-load latency no longer sits in front of every tile's arithmetic.  Needs unmasked complex
    // loads of tiles whose lanes are contiguous in memory (the planner checks).
    TM_PIPE_C2C = 4,
    // DCT-II of real rows through ONE half-length complex transform (Makhoul): the even / reversed-odd input
    // permutation is the load pattern, X[k] = Re(w_k V[k]), X[N-k] = -Im(w_k V[k]) the store (dct.rs:523-559)
    TM_FAST_DCT2 = 5,
    // the inverse packing: V[k] = conj(w_k)(X[k] - i X[N-k]) -> c2r pre-twiddle -> half-length transform -> scatter:
    // y[i] = scale * 2 * (scale_dc * X[0] / 2 + sum_{k>=1} X[k] cos(pi k (i + 1/2) / N))   (dct.rs:563-684)
    TM_FAST_DCT3 = 6,
    // DCT-IV / DST-IV of N = 2L reals through ONE L-point complex transform (dct.rs:688-720, dst.rs:630-667):
    //   z[j] = (x[2j] + i x[N-1-2j]) exp(-i pi (4j+1) / (4N)),  Z = FFT_L(z),  y[k] = Z[k] exp(-i pi k / N),
    //   X[2k] = Re y[k],  X[N-1-2k] = -Im y[k];  the sine transform is the cosine transform of the reversed input with
    //   (-1)^k on the output.  NOT YET RUN ON A GPU (written after the round's GPU budget was spent): the planner only
    //   takes it with SFC_DCT4_FUSED=1, and tests/test_gpu_experimental.py is the parity check to run first.
    TM_FAST_DCT4 = 7,
    // The middle pass of the three-pass plan for large 2-D transforms (tools/fft2_three_pass_emulation.py):
    // a strided tile of L = 16*LB points whose stages skip the twiddles after the FIRST radix-16 stage, which makes it the
    // 16 x LB two-dimensional transform of the blocks e = LB*r_lo + c_hi; output q = k2 + 16*kc1 is multiplied by
    // aux_out[kc1 * c_rest] (W_C^j) and stored at k2*mid_es + kc1*mid_ls (the two mid_* strides are free here: not a double
    // kernel).  NOT YET RUN ON A GPU: planner knob SFC_FFT2_TILE2D=1, parity check in tests/test_gpu_experimental.py.
    TM_FAST_2D = 8,
    // TM_FAST_C2C made persistent with a LATE prefetch that costs no shared memory: once the last exchange of a tile has
    // been read back into registers the exchange buffer is idle for the rest of the tile (last radix stage, store
    // operators, stores), so the TMA unit lands the NEXT tile in it (cp.async.bulk + mbarrier) during that time.  The
    // next tile then starts with its data already on chip: the global-load latency of every tile but the first hides
    // behind the previous tile's tail, with the full-size exchange buffer and the same number of CTAs per SM as the
    // plain flavour.  Contiguous-row tiles (1-D bulk copies) and, through a tensor map, strided column tiles.
    TM_PIPE_LATE = 9,
};

// ---- TMA / mbarrier primitives (PTX) -----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// one box of a 4-D tiled tensor map -> shared memory, completion on an mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
        : "memory");
}

// blockIdx-level index -> (tile, batch)
__device__ __forceinline__ void decode_block(const PassParams& p, uint32_t blk, uint32_t& tile, uint32_t& batch) {
    if (p.nbatch_fast) {
        const uint32_t per = p.nbatch_fast << p.tile_group_shift;
        const uint32_t th = blk / per, r = blk - th * per;
        batch = r >> p.tile_group_shift;
        tile = (th << p.tile_group_shift) + (r & ((1u << p.tile_group_shift) - 1u));
    } else if (p.win_len) {
        const uint32_t row = blk / p.win_len;
        tile = row * p.win_row_tiles + p.win_first + (blk - row * p.win_len);
        batch = 0;
    } else {
        tile = blk % p.tiles_per_batch;
        batch = blk / p.tiles_per_batch;
    }
}

// warp 0 of a pipelined CTA: land tile `blk` (TL lanes x L complex, dense) in shared memory
template <typename T, int L, int TL>
__device__ __forceinline__ void pipe_issue(const PassParams& p, uint32_t blk, uint32_t land, uint32_t bar) {
    const int lane_id = threadIdx.x & 31;
    uint32_t tile, batch;
    decode_block(p, blk, tile, batch);
    if (lane_id == 0) mbar_expect_tx(bar, (uint32_t)(TL * L * sizeof(Cx<T>)));
    __syncwarp();
    const Cx<T>* base = reinterpret_cast<const Cx<T>*>(p.in.ptr) + (int64_t)batch * p.in.batch_stride;
    if (p.map_in == MAP_COL && (p.flags & F_TMAP_IN)) {
        // strided tile through the tensor map: boxes of [tmap_box_rows elements][TL lanes], dense in shared memory
        const uint32_t lane0 = tile * TL;
        int c0 = (int)lane0, c2 = 0;
        if (p.tmap_split) {
            const uint32_t lo = lane0 / p.inner_count;
            c0 = (int)(lane0 - lo * p.inner_count);
            c2 = (int)lo;
        }
        const int rows = p.tmap_box_rows;
        for (int r = lane_id * rows; r < L; r += 32 * rows)
            tma_load_4d(land + (uint32_t)(r * TL * sizeof(Cx<T>)), p.tmap_in, 2 * c0, r, c2, (int)batch, bar);
    } else if (p.map_in == MAP_COL) {
        // element rows of TL adjacent lanes: L copies of TL * sizeof(cx) bytes
        const uint32_t lane0 = tile * TL;
        const uint32_t lo = lane0 / p.inner_count, li = lane0 - lo * p.inner_count;
        const Cx<T>* src = base + (int64_t)lo * p.in.outer_stride + (int64_t)li * p.in.inner_stride;
        for (int e = lane_id; e < L; e += 32)
            bulk_g2s(land + (uint32_t)(e * TL * sizeof(Cx<T>)), src + (int64_t)e * p.in.elem_stride,
                     (uint32_t)(TL * sizeof(Cx<T>)), bar);
    } else {
        // contiguous rows: one copy per lane
        for (int t = lane_id; t < TL; t += 32) {
            const uint32_t lane = tile * TL + (uint32_t)t;
            const uint32_t lo = lane / p.inner_count, li = lane - lo * p.inner_count;
            const Cx<T>* src = base + (int64_t)lo * p.in.outer_stride + (int64_t)li * p.in.inner_stride;
            bulk_g2s(land + (uint32_t)(t * L * sizeof(Cx<T>)), src, (uint32_t)(L * sizeof(Cx<T>)), bar);
        }
    }
}

template <typename T, int L, int TL>
__device__ __forceinline__ void late_issue(const LateIssue& h) {
    if (threadIdx.x < 32 && h.next_blk < h.p->total_tiles) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy reads of the buffer before the async-proxy writes
        pipe_issue<T, L, TL>(*h.p, h.next_blk, h.land, h.bar);
    }
}

// Z[k] = (X[k] + conj X[L-k]) + i*conj(W_2L^k)*(X[k] - conj X[L-k]) from a staged row
template <typename T, typename C>
__device__ __forceinline__ void c2r_pretwiddle(Cx<T> (&a)[C::E], const Cx<T>* row, const Cx<T>* rtw, int i0) {
    constexpr int E = C::E, TPL = C::TPL, L = C::L;
    if constexpr (E >= 8) {
        const Cx<T> wi = rtw[i0];
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const int k = i0 + m * TPL;
            Cx<T> xk = row[k];
            Cx<T> xp = row[L - k];
            if (k == 0) {  // imag of DC / Nyquist never reaches the real output
                xk.y = 0;
                xp.y = 0;
            }
            const Cx<T> A = {xk.x + xp.x, xk.y - xp.y};
            const Cx<T> D = {xk.x - xp.x, xk.y + xp.y};
            const Cx<T> w = cmul(wi, w32<T>(m * (16 / E)));
            const Cx<T> Bc = cmulc(D, w);        // conj(W) * D
            a[m] = {A.x - Bc.y, A.y + Bc.x};     // A + i*Bc
        }
    }
}

// one tile: load -> transform -> store.  Pipelined flavour: `land` / `bar` are the landing buffer and its
// mbarrier, `parity` the phase to wait for, `next_blk` the tile to prefetch once this one is in registers and
// `bar_next` the barrier that prefetch completes on (the other thread group's in the group-pipelined flavour).
template <typename T, int L, int TL, bool DOUBLE, int EMAX, int MODE, int GROUPS>
__device__ __forceinline__ void tile_body(const PassParams& p, const uint32_t blk, Cx<T>* sm, const Cx<T>* land,
                                          uint32_t bar, uint32_t parity, uint32_t next_blk, uint32_t bar_next) {
    constexpr bool PIPE = MODE == TM_PIPE_C2C;
    constexpr bool LATE = MODE == TM_PIPE_LATE;  // data of this tile was landed in the exchange buffer during the previous tile's tail
    constexpr bool GP = PIPE && GROUPS == 2;  // group-pipelined: each thread group walks over its own tiles of TLG lanes
    using C = TileCfg<T, L, TL, EMAX, GROUPS, PIPE && GROUPS == 1>;
    constexpr int LT = GP ? C::TLG : TL;      // lanes per scheduled tile
    using cx = Cx<T>;
    constexpr int E = C::E, TPL = C::TPL, LP = C::LP;
    constexpr bool FAST = MODE != TM_GENERIC;

    const int tid = threadIdx.x;
    // tile-major order (batch index fastest) lets consecutive CTAs reuse the same rows of the
    // per-plan tables (Bluestein chirp / kernel spectrum) out of L2; groups of 2^shift adjacent
    // tiles stay adjacent in time so that their 64..128 B row segments merge into whole DRAM bursts
    uint32_t tile, batch;
    decode_block(p, blk, tile, batch);

    int t0, i0, t1, i1;
    const int grp = GROUPS == 1 ? 0 : tid / C::NTG;  // warp-uniform
    {
        const int tg = GROUPS == 1 ? tid : tid % C::NTG;
        map_thread<C::TLG, TPL>(p.map_in, tg, t0, i0);
        map_thread<C::TLG, TPL>(p.map_out, tg, t1, i1);
        t0 += grp * C::TLG;
        t1 += grp * C::TLG;
    }

    cx a[E];
    bool staged = false;
#ifdef SFC_PHASE_TIMING
    long long phase_t_ = clock64();
    if (p.dbg && threadIdx.x == 0) atomicAdd(p.dbg + 14, 1ull);
#endif

    // ------------------------------ load ------------------------------------
    {
        const uint32_t lane = tile * LT + (uint32_t)(GP ? t0 - grp * C::TLG : t0);
        const bool valid = lane < p.nlanes;
        const uint32_t lo = lane / p.inner_count;
        const uint32_t li = lane - lo * p.inner_count;
        const int64_t off = (int64_t)batch * p.in.batch_stride + (int64_t)lo * p.in.outer_stride +
                            (int64_t)li * p.in.inner_stride;
        const int64_t pos0 = (int64_t)lo * p.in.pos_ls;
        constexpr bool C2R_MIRROR = (MODE == TM_FAST_C2R) && E == 16 && L >= 128 && (ilog2(L) % 4 == 3);
        if constexpr (C2R_MIRROR) {
            // thread i loads butterfly i (X[i + r*SL]) and its mirror (X[SL - i + r*SL]) of the leading
            // radix-8 stage: X[k] and X[L-k] are then both in its registers and the pre-twiddle
            //   Z[k] = A + B,  Z[L-k] = conj(A - B),  A = X[k] + conj X[L-k],  B = i conj(W^k)(X[k] - conj X[L-k])
            // needs neither staging nor a barrier.
            constexpr int SL = L / 8;
            const cx* __restrict__ src = reinterpret_cast<const cx*>(p.in.ptr) + off;
            const int j2 = i0 == 0 ? SL / 2 : SL - i0;
            cx u[8], v[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                u[r] = src[i0 + r * SL];
                v[r] = src[j2 + r * SL];
            }
            auto pre = [&](cx xk, cx xp, cx w, cx& zk, cx& zp) {
                const cx A = {xk.x + xp.x, xk.y - xp.y};
                const cx D = {xk.x - xp.x, xk.y + xp.y};
                const cx Bc = cmulc(D, w);            // conj(W) * D
                zk = {A.x - Bc.y, A.y + Bc.x};        // A + i*Bc
                zp = {A.x + Bc.y, -(A.y - Bc.x)};     // conj(A - i*Bc)
            };
            if (i0 != 0) {
                const cx wi = reinterpret_cast<const cx*>(p.rtw)[i0];
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    pre(u[r], v[7 - r], cmul(wi, w32<T>(2 * r)), a[2 * r], a[2 * (7 - r) + 1]);
            } else {
                cx x0 = u[0], xl = src[L], dummy;
                x0.y = 0;  // imag of DC / Nyquist never reaches the real output
                xl.y = 0;
                pre(x0, xl, w32<T>(0), a[0], dummy);
#pragma unroll
                for (int r = 1; r < 4; ++r) pre(u[r], u[8 - r], w32<T>(2 * r), a[2 * r], a[2 * (8 - r)]);
                pre(u[4], u[4], w32<T>(8), a[8], dummy);
#pragma unroll
                for (int r = 0; r < 4; ++r) pre(v[r], v[7 - r], w32<T>(2 * r + 1), a[2 * r + 1], a[2 * (7 - r) + 1]);
            }
        } else if constexpr (MODE == TM_FAST_C2R) {
            // stage the L+1 Hermitian inputs of each lane (unit stride), then pre-twiddle.
            // (Measured alternative: every thread loading X[k] and its mirror X[L-k] straight from
            // global memory, no staging — 71 % -> 67 % of HBM peak, so staging stays.)
            const cx* __restrict__ src = reinterpret_cast<const cx*>(p.in.ptr) + off + i0;
            cx* row = sm + t0 * LP;
#pragma unroll
            for (int m = 0; m < E; ++m) row[i0 + m * TPL] = src[m * TPL];
            if (i0 == 0) row[L] = src[L];
            C::sync(grp);
            c2r_pretwiddle<T, C>(a, row, reinterpret_cast<const cx*>(p.rtw), i0);
            staged = true;
        } else if (FAST && !PIPE && p.ld_op == LD_SPLIT2) {
            // One 2L-point row is shared by two lanes (usually two CTAs): a radix-2 decimation-in-
            // frequency stage folded into the load.  Both lanes read the whole row (the second read
            // comes from L2); lane parity 0 transforms x[j] + x[j+L] -> even bins, parity 1 transforms
            // (x[j] - x[j+L]) W_2L^j -> odd bins.  Two 64 KiB tiles per SM instead of one 128 KiB tile.
            if constexpr (E == 16) {
                const cx* __restrict__ src = reinterpret_cast<const cx*>(p.in.ptr) + off + i0;
                const T sgn = li ? (T)-1 : (T)1;
                const T cj = (p.flags & F_CONJ_LD_PRE) ? (T)-1 : (T)1;
#pragma unroll
                for (int m = 0; m < E; ++m) {
                    const cx u = src[m * TPL], v = src[m * TPL + L];
                    a[m] = {fma(sgn, v.x, u.x), cj * fma(sgn, v.y, u.y)};
                }
                if (li) {
                    const cx wi = reinterpret_cast<const cx*>(p.rtw)[i0];
#pragma unroll
                    for (int m = 0; m < E; ++m) a[m] = cmul(a[m], cmul(wi, w32<T>(m)));
                }
            }
        } else if constexpr (MODE == TM_FAST_DCT2) {
            // v[p] = x[2p] (p < N/2), x[2N-2p-1] (p >= N/2), packed as z[j] = v[2j] + i v[2j+1]; N = 2L reals per lane.
            // The I/O descriptors of the DCT flavours count REAL elements (rows: elem_stride 1; column tiles: the
            // stride of the transformed axis, adjacent lanes adjacent in memory).
            if constexpr (E == 16) {
                const T* __restrict__ xr = reinterpret_cast<const T*>(p.in.ptr) + off;
                const int64_t es = p.in.elem_stride;
#pragma unroll
                for (int m = 0; m < E; ++m) {
                    const int j = i0 + m * TPL;
                    if (m < E / 2) a[m] = {xr[(4 * j) * es], xr[(4 * j + 2) * es]};
                    else a[m] = {xr[(4 * L - 4 * j - 1) * es], xr[(4 * L - 4 * j - 3) * es]};
                }
                if (p.flags & F_TRIG_SINE) {  // (-1)^i x[i]: the second half holds the odd-indexed samples
#pragma unroll
                    for (int m = E / 2; m < E; ++m) a[m] = {-a[m].x, -a[m].y};
                }
            }
        } else if constexpr (MODE == TM_FAST_DCT3) {
            if constexpr (E == 16) {
                // stage V[k] = conj(w_k) (X[k] - i X[N-k]), k <= L (X[N] = 0), then the c2r pre-twiddle
                constexpr int N = 2 * L;
                const T* __restrict__ xr = reinterpret_cast<const T*>(p.in.ptr) + off;
                const int64_t es = p.in.elem_stride;
                const cx* __restrict__ om = reinterpret_cast<const cx*>(p.aux_out);
                const bool rev = (p.flags & F_TRIG_SINE) != 0;  // DST-III reads its input reversed
                auto X = [&](int k) { return rev ? xr[(N - 1 - k) * es] : xr[k * es]; };
                cx* row = sm + t0 * LP;
#pragma unroll
                for (int m = 0; m < E; ++m) {
                    const int k = i0 + m * TPL;
                    cx u = {X(k), k == 0 ? (T)0 : -X(N - k)};
                    if (k == 0) u.x *= (T)p.scale_dc;
                    row[k] = cmulc(u, om[k]);
                }
                if (i0 == 0) row[L] = cmulc(cx{X(L), -X(L)}, om[L]);
                C::sync(grp);
                c2r_pretwiddle<T, C>(a, row, reinterpret_cast<const cx*>(p.rtw), i0);
                staged = true;
            }
        } else if constexpr (MODE == TM_FAST_DCT4) {
            if constexpr (E == 16) {
                constexpr int N = 2 * L;
                const T* __restrict__ xr = reinterpret_cast<const T*>(p.in.ptr) + off;
                const int64_t es = p.in.elem_stride;
                const cx* __restrict__ pre = reinterpret_cast<const cx*>(p.aux_in);  // exp(-i pi (4j+1) / (4N)), j < L
                const bool sine = (p.flags & F_TRIG_SINE) != 0;
#pragma unroll
                for (int m = 0; m < E; ++m) {
                    const int j = i0 + m * TPL;
                    const T ev = xr[(2 * j) * es], od = xr[(N - 1 - 2 * j) * es];
                    a[m] = cmul(sine ? cx{od, ev} : cx{ev, od}, pre[j]);
                }
            }
        } else if constexpr (LATE) {
            mbar_wait(bar, parity);
            const cx* src = (p.map_in == MAP_COL) ? land + (i0 * TL + t0) : land + (t0 * L + i0);
            const int step = (p.map_in == MAP_COL) ? TPL * TL : TPL;
#pragma unroll
            for (int m = 0; m < E; ++m) a[m] = src[m * step];
            // (the barrier that frees the buffer for the first exchange is the one run_stages<FIRST = false> starts with)
            if (p.flags & F_CONJ_LD_PRE) {
#pragma unroll
                for (int m = 0; m < E; ++m) a[m].y = -a[m].y;
            }
            if (p.ld_op == LD_C_MUL && (p.flags & F_CHIRP_GEN)) {
                chirp_apply<T, E>(a, p, (int64_t)i0 * p.in.pos_es + pos0, (int64_t)TPL * p.in.pos_es, p.chirp_q_in);
            } else if (p.ld_op == LD_C_MUL) {
                const cx* __restrict__ aux = reinterpret_cast<const cx*>(p.aux_in);
#pragma unroll
                for (int m = 0; m < E; ++m) a[m] = cmul(a[m], aux[(int64_t)(i0 + m * TPL) * p.in.pos_es + pos0]);
            }
        } else if constexpr (PIPE) {
            // the tile was landed in shared memory by the TMA unit while the previous one was transformed
            mbar_wait(bar, parity);
            const int tl = GP ? t0 - grp * C::TLG : t0;  // lane inside the landed tile
            const cx* src = (p.map_in == MAP_COL) ? land + (i0 * LT + tl) : land + (tl * L + i0);
            const int step = (p.map_in == MAP_COL) ? TPL * LT : TPL;
#pragma unroll
            for (int m = 0; m < E; ++m) a[m] = src[m * step];
            C::sync(grp);  // everybody (of this group) holds its elements: the landing buffer is free again
            if ((GP ? tid - grp * C::NTG : tid) < 32 && next_blk < p.total_tiles) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                // group-pipelined: the next tile belongs to the OTHER group and completes on its barrier
                pipe_issue<T, L, LT>(p, next_blk, smem_u32(land), bar_next);
            }
            if (p.flags & F_CONJ_LD_PRE) {
#pragma unroll
                for (int m = 0; m < E; ++m) a[m].y = -a[m].y;
            }
            if (p.ld_op == LD_C_MUL && (p.flags & F_CHIRP_GEN)) {
                chirp_apply<T, E>(a, p, (int64_t)i0 * p.in.pos_es + pos0, (int64_t)TPL * p.in.pos_es, p.chirp_q_in);
            } else if (p.ld_op == LD_C_MUL) {
                const cx* __restrict__ aux = reinterpret_cast<const cx*>(p.aux_in);
#pragma unroll
                for (int m = 0; m < E; ++m) a[m] = cmul(a[m], aux[(int64_t)(i0 + m * TPL) * p.in.pos_es + pos0]);
            }
        } else if constexpr (FAST) {
            const cx* __restrict__ src =
                reinterpret_cast<const cx*>(p.in.ptr) + off + (int64_t)i0 * p.in.elem_stride;
            const int64_t step = (int64_t)TPL * p.in.elem_stride;
            // optional zero padding: element e is real data iff e*pos_es + pos0 < len, i.e. e < elim
            int elim = L;
            bool staged_in = false;
            if constexpr (L <= 64 && MODE == TM_FAST_C2C) staged_in = (p.flags & F_STAGE_IN) != 0;
            if (staged_in) {
                if constexpr (L <= 64 && MODE == TM_FAST_C2C) {
                    // staging pitch: a quarter warp (8 threads x 16 B) reads TPL adjacent elements of 8/TPL rows
                    constexpr int SP = L + (TPL < 8 ? TPL : 8);
                    static_assert((size_t)TL * SP <= (size_t)TL * LP, "staging rows must fit the exchange buffer");
                    const cx* __restrict__ tb = reinterpret_cast<const cx*>(p.in.ptr) + (int64_t)batch * p.in.batch_stride +
                                                (int64_t)tile * (TL * L);
#pragma unroll
                    for (int k = 0; k < E; ++k) {
                        const int f = tid + k * C::NT;
                        sm[(f / L) * SP + (f % L)] = tb[f];
                    }
                    C::sync(grp);
#pragma unroll
                    for (int m = 0; m < E; ++m) a[m] = sm[t0 * SP + i0 + m * TPL];
                    C::sync(grp);  // the exchange (or the staged store) reuses the buffer
                }
            } else if (!(p.flags & F_IN_NOMASK)) {
                const int64_t rem = p.in.len - pos0;
                elim = rem <= 0 ? 0 : (int)min((uint32_t)L, ((uint32_t)rem + (uint32_t)p.in.pos_es - 1u) / (uint32_t)p.in.pos_es);
#pragma unroll
                for (int m = 0; m < E; ++m) {
                    cx v = {(T)0, (T)0};
                    if (i0 + m * TPL < elim) v = ld_pol<SFC_LDPOL_DATA>(src + m * step);
                    a[m] = v;
                }
            } else {
#pragma unroll
                for (int m = 0; m < E; ++m) a[m] = ld_pol<SFC_LDPOL_DATA>(src + m * step);
            }
            SFC_PHASE(0);          // address set-up + load issue
            SFC_PHASE_FORCE(a);
            SFC_PHASE(1);          // waiting for the tile data
            if (p.flags & F_CONJ_LD_PRE) {
#pragma unroll
                for (int m = 0; m < E; ++m) a[m].y = -a[m].y;
            }
            if (p.ld_op == LD_C_MUL && (p.flags & F_CHIRP_GEN)) {
                // zero-padded elements stay zero whatever they are multiplied by: no predicate
                chirp_apply<T, E>(a, p, (int64_t)i0 * p.in.pos_es + pos0, (int64_t)TPL * p.in.pos_es, p.chirp_q_in);
            } else if (p.ld_op == LD_C_MUL) {
                const cx* __restrict__ aux = reinterpret_cast<const cx*>(p.aux_in);
#pragma unroll
                for (int m = 0; m < E; ++m)
                    if (i0 + m * TPL < elim)
                        a[m] = cmul(a[m], ld_pol<SFC_LDPOL_AUX>(aux + ((int64_t)(i0 + m * TPL) * p.in.pos_es + pos0)));
            }
        } else if (p.ld_op == LD_C2R) {
            const cx* __restrict__ src = reinterpret_cast<const cx*>(p.in.ptr) + off;
            cx* row = sm + t0 * LP;
#pragma unroll
            for (int m = 0; m < E; ++m) {
                const int e = i0 + m * TPL;
                cx v = {0, 0};
                if (valid && (int64_t)e < p.in.len) v = src[(int64_t)e * p.in.elem_stride];
                row[e] = v;
            }
            if (i0 == 0) {
                cx v = {0, 0};
                if (valid && (int64_t)L < p.in.len) v = src[(int64_t)L * p.in.elem_stride];
                row[L] = v;
            }
            C::sync(grp);
            c2r_pretwiddle<T, C>(a, row, reinterpret_cast<const cx*>(p.rtw), i0);
            staged = true;
        } else {
            const bool is_real = (p.ld_op == LD_R) || (p.ld_op == LD_R_MUL);
            const bool has_mul = (p.ld_op == LD_C_MUL) || (p.ld_op == LD_R_MUL);
            if (is_real) {
                const T* __restrict__ src = reinterpret_cast<const T*>(p.in.ptr) + off;
#pragma unroll
                for (int m = 0; m < E; ++m) {
                    const int e = i0 + m * TPL;
                    const int64_t pos = (int64_t)e * p.in.pos_es + pos0;
                    T r = 0;
                    if (valid && pos < p.in.len) r = src[(int64_t)e * p.in.elem_stride];
                    a[m] = {r, (T)0};
                }
            } else {
                const cx* __restrict__ src = reinterpret_cast<const cx*>(p.in.ptr) + off;
#pragma unroll
                for (int m = 0; m < E; ++m) {
                    const int e = i0 + m * TPL;
                    const int64_t pos = (int64_t)e * p.in.pos_es + pos0;
                    cx v = {0, 0};
                    if (valid && pos < p.in.len) v = src[(int64_t)e * p.in.elem_stride];
                    a[m] = v;
                }
            }
            if (p.flags & F_CONJ_LD_PRE) {
#pragma unroll
                for (int m = 0; m < E; ++m) a[m].y = -a[m].y;
            }
            if (has_mul && (p.flags & F_CHIRP_GEN)) {
                chirp_apply<T, E>(a, p, (int64_t)i0 * p.in.pos_es + pos0, (int64_t)TPL * p.in.pos_es, p.chirp_q_in);
            } else if (has_mul) {
                const cx* __restrict__ aux = reinterpret_cast<const cx*>(p.aux_in);
#pragma unroll
                for (int m = 0; m < E; ++m) {
                    const int e = i0 + m * TPL;
                    const int64_t pos = (int64_t)e * p.in.pos_es + pos0;
                    if (pos < p.in.len) a[m] = cmul(a[m], aux[pos]);
                }
            }
        }
        if (p.flags & F_LD_TW)
            fourstep_twiddle<T, E, TPL>(a, p.ld_tw_lo, p.ld_tw_hi, p.ld_tw_shift, (p.flags & F_LD_TW_CONJ) != 0, i0, lo);
        if (MODE == TM_FAST_C2R || MODE == TM_FAST_DCT3 || (p.flags & F_CONJ_LD_POST)) {
#pragma unroll
            for (int m = 0; m < E; ++m) a[m].y = -a[m].y;
        }
    }

    SFC_PHASE_FORCE(a);
    SFC_PHASE(2);  // load operators (chirp multiply, ...)
    // ---------------------------- transform ---------------------------------
    const cx* __restrict__ tw = reinterpret_cast<const cx*>(p.tw);
    // last radix 8 with 16 points per thread = two butterflies per thread in the last stage
    constexpr bool R2C_MIRROR =
        (MODE == TM_FAST_R2C) && E == 16 && L >= 128 && (ilog2(L) % 4 == 3) && sizeof(T) == 8;  // f32: measured slower
    if constexpr (R2C_MIRROR) {
        run_stages<T, C, 1, true, true>(a, sm, tw, t0, i0, t1, i1, grp);
    } else if constexpr ((MODE == TM_FAST_C2R) && E == 16 && L >= 128 && (ilog2(L) % 4 == 3)) {
        run_stages<T, C, 1, true, false, true>(a, sm, tw, t0, i0, t1, i1, grp);
    } else if constexpr (MODE == TM_FAST_C2R || MODE == TM_FAST_DCT3) {
        run_stages<T, C, 1, false>(a, sm, tw, t0, i0, t1, i1, grp);
    } else if constexpr (MODE == TM_FAST_2D) {
        run_stages<T, C, 1, true, false, false, true>(a, sm, tw, t0, i0, t1, i1, grp);
    } else if constexpr (LATE) {
        static_assert(!DOUBLE && C::E < C::L && GROUPS == 1, "late prefetch: plain multi-stage complex tiles");
        const LateIssue hook{&p, next_blk, bar_next, smem_u32(sm)};
        run_stages<T, C, 1, false, false, false, false, true>(a, sm, tw, t0, i0, t1, i1, grp, &hook);
    } else if constexpr (FAST) {
        run_stages<T, C, 1, true>(a, sm, tw, t0, i0, t1, i1, grp);
    } else {
        if (staged)
            run_stages<T, C, 1, false>(a, sm, tw, t0, i0, t1, i1, grp);
        else
            run_stages<T, C, 1, true>(a, sm, tw, t0, i0, t1, i1, grp);
    }

    if constexpr (C::E == C::L) {
        // single-stage tiles never touch shared memory: remap explicitly if the
        // store wants the other thread->lane mapping
        if (p.map_in != p.map_out) {
            C::sync(grp);
#pragma unroll
            for (int m = 0; m < E; ++m) sm[t0 * LP + m] = a[m];
            C::sync(grp);
#pragma unroll
            for (int m = 0; m < E; ++m) a[m] = sm[t1 * LP + m];
        }
    }

    SFC_PHASE_FORCE(a);
    SFC_PHASE(3);  // first transform
    const uint32_t lane = tile * LT + (uint32_t)(GP ? t1 - grp * C::TLG : t1);
    const bool valid = FAST ? true : (lane < p.nlanes);
    const uint32_t lo = lane / p.inner_count;
    const uint32_t li = lane - lo * p.inner_count;

    if constexpr (DOUBLE) {
        // forward transform -> pointwise table -> inverse transform, all on chip
        const cx* __restrict__ mid = reinterpret_cast<const cx*>(p.mid);
        const int64_t mid0 = (int64_t)lo * p.mid_ls + (int64_t)li * p.mid_is;
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const int e = i1 + m * TPL;
            cx w = mid[(int64_t)e * p.mid_es + mid0];
            a[m] = cconjf(cmul(a[m], w));
        }
        run_stages<T, C, 1, false>(a, sm, tw, t1, i1, t1, i1, grp);
#pragma unroll
        for (int m = 0; m < E; ++m) a[m].y = -a[m].y;
    }

    SFC_PHASE_FORCE(a);
    SFC_PHASE(4);  // pointwise table + second transform (double kernels)
    // ------------------------------ store -----------------------------------
    const int64_t off = (int64_t)batch * p.out.batch_stride + (int64_t)lo * p.out.outer_stride +
                        (int64_t)li * p.out.inner_stride;
    const int64_t pos0 = (int64_t)lo * p.out.pos_ls;
    const T scale = (T)p.scale;

    if (MODE == TM_FAST_C2R || MODE == TM_FAST_DCT3 || (p.flags & F_CONJ_ST_PRE)) {
#pragma unroll
        for (int m = 0; m < E; ++m) a[m].y = -a[m].y;
    }

    if constexpr (MODE == TM_FAST_2D) {
        if constexpr (E == 16 && L >= 32) {
            // q = i1 + m*TPL = k2 + 16*kc1 (the first stage's output digit is the least significant one)
            const cx* __restrict__ tw2 = reinterpret_cast<const cx*>(p.aux_out);
            cx* __restrict__ dst = reinterpret_cast<cx*>(p.out.ptr) + off;
#pragma unroll
            for (int m = 0; m < E; ++m) {
                const int q = i1 + m * TPL;
                const int k2 = q & 15, kc1 = q >> 4;
                const cx v = cmul(a[m], tw2[(int64_t)kc1 * li]);
                dst[(int64_t)k2 * p.mid_es + (int64_t)kc1 * p.mid_ls] = {v.x * scale, v.y * scale};
            }
        }
        return;
    } else if constexpr (MODE == TM_FAST_DCT4) {
        if constexpr (E == 16) {
            constexpr int N = 2 * L;
            T* __restrict__ dst = reinterpret_cast<T*>(p.out.ptr) + off;
            const int64_t oes = p.out.elem_stride;
            const cx* __restrict__ post = reinterpret_cast<const cx*>(p.aux_out);  // exp(-i pi k / N), k < L
            const T so = (p.flags & F_TRIG_SINE) ? scale : -scale;  // odd outputs: -Im y (cosine), +Im y (sine: (-1)^(N-1-2k) = -1)
#pragma unroll
            for (int m = 0; m < E; ++m) {
                const int k = i1 + m * TPL;
                const cx y = cmul(a[m], post[k]);
                dst[(2 * k) * oes] = y.x * scale;
                dst[(N - 1 - 2 * k) * oes] = y.y * so;
            }
        }
        return;
    } else if constexpr (MODE == TM_FAST_DCT2) {
        if constexpr (E == 16) {
            // a[m] = Z[i1 + m*TPL] of the packed half-length transform -> V[k], V[L-k] of the real transform as in
            // the r2c post-pass, then four real outputs per pair
            cx* row = sm + t1 * LP;
            C::sync(grp);
#pragma unroll
            for (int m = 0; m < E; ++m) row[i1 + m * TPL] = a[m];
            C::sync(grp);
            constexpr int N = 2 * L;
            // DST-II: output k of the cosine transform is output N-1-k of the sine transform
            const bool rev = (p.flags & F_TRIG_SINE) != 0;
            const int64_t oes = p.out.elem_stride;
            T* __restrict__ dst = reinterpret_cast<T*>(p.out.ptr) + off + (rev ? (N - 1) * oes : 0);
            const int64_t ds = rev ? -oes : oes;
            const cx* __restrict__ om = reinterpret_cast<const cx*>(p.aux_out);  // w_k = exp(-i pi k / (2N)), k <= L
            const cx wi = reinterpret_cast<const cx*>(p.rtw)[i1];
            const T h = (T)0.5 * scale;
#pragma unroll
            for (int m = 0; m < E / 2; ++m) {
                const int k = i1 + m * TPL;
                const cx zk = a[m];
                const cx zp = row[(L - k) & (L - 1)];
                const cx A = {zk.x + zp.x, zk.y - zp.y};
                const cx B = {zk.x - zp.x, zk.y + zp.y};
                const cx Cw = cmul(B, cmul(wi, w32<T>(m * (16 / E))));
                const cx vk = {(A.x + Cw.y) * h, (A.y - Cw.x) * h};       // V[k]
                const cx vq = {(A.x - Cw.y) * h, -((A.y + Cw.x) * h)};    // V[L-k]
                const cx tk = cmul(vk, om[k]);
                const cx tq = cmul(vq, om[L - k]);
                if (k == 0) {
                    dst[0] = tk.x * (T)p.scale_dc;   // X[0] (the "ortho" 1/sqrt(2), dct.rs:552-553)
                    dst[ds * L] = tq.x;              // X[N/2]
                } else {
                    dst[ds * k] = tk.x;
                    dst[ds * (N - k)] = -tk.y;
                    dst[ds * (L - k)] = tq.x;
                    dst[ds * (L + k)] = -tq.y;
                }
            }
            if (i1 == 0) {
                const cx z = a[E / 2];  // Z[L/2]: V[L/2] = conj(Z[L/2])
                const cx t = cmul(cx{z.x * scale, -(z.y * scale)}, om[L / 2]);
                dst[ds * (L / 2)] = t.x;
                dst[ds * (N - L / 2)] = -t.y;
            }
        }
        return;
    } else if constexpr (MODE == TM_FAST_DCT3) {
        if constexpr (E == 16) {
            // a[m] = (v[2j], v[2j+1]), j = i1 + m*TPL: undo the even / reversed-odd permutation on the way out
            T* __restrict__ dst = reinterpret_cast<T*>(p.out.ptr) + off;
            const int64_t oes = p.out.elem_stride;
            const T sg = (p.flags & F_TRIG_SINE) ? -scale : scale;  // DST-III: (-1)^i on the odd-indexed outputs
#pragma unroll
            for (int m = 0; m < E; ++m) {
                const int j = i1 + m * TPL;
                if (m < E / 2) {
                    dst[(4 * j) * oes] = a[m].x * scale;
                    dst[(4 * j + 2) * oes] = a[m].y * scale;
                } else {
                    dst[(4 * L - 4 * j - 1) * oes] = a[m].x * sg;
                    dst[(4 * L - 4 * j - 3) * oes] = a[m].y * sg;
                }
            }
        }
        return;
    } else if constexpr (R2C_MIRROR) {
        // a[2*kk] = Z[i1 + kk*SL], a[2*kk+1] = Z[j2 + kk*SL] with j2 the mirror butterfly:
        // Z[L-(i1 + kk*SL)] = a[2*(7-kk)+1] — the whole Hermitian post-pass stays in registers.
        constexpr int SL = L / 8;
        cx* __restrict__ dst = reinterpret_cast<cx*>(p.out.ptr) + off;
        const T h = (T)0.5 * scale;
        auto emit = [&](cx zk, cx zp, cx w, int k) {
            const cx A = {zk.x + zp.x, zk.y - zp.y};
            const cx B = {zk.x - zp.x, zk.y + zp.y};
            const cx Cw = cmul(B, w);
            dst[k] = {(A.x + Cw.y) * h, (A.y - Cw.x) * h};
            dst[L - k] = {(A.x - Cw.y) * h, -((A.y + Cw.x) * h)};
        };
        if (i1 != 0) {
            const cx wi = reinterpret_cast<const cx*>(p.rtw)[i1];
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
                emit(a[2 * kk], a[2 * (7 - kk) + 1], cmul(wi, w32<T>(2 * kk)), i1 + kk * SL);
        } else {
            // butterfly 0 pairs kk <-> 8-kk (kk = 0: DC / Nyquist, kk = 4: self), butterfly SL/2 pairs kk <-> 7-kk
            emit(a[0], a[0], w32<T>(0), 0);
#pragma unroll
            for (int kk = 1; kk < 4; ++kk) emit(a[2 * kk], a[2 * (8 - kk)], w32<T>(2 * kk), kk * SL);
            dst[L / 2] = {a[8].x * scale, -(a[8].y * scale)};
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
                emit(a[2 * kk + 1], a[2 * (7 - kk) + 1], w32<T>(2 * kk + 1), SL / 2 + kk * SL);
        }
        return;
    } else if (MODE == TM_FAST_R2C || (MODE == TM_GENERIC && p.st_op == ST_R2C)) {
        if constexpr (E >= 8 && E != L && (MODE == TM_FAST_R2C || MODE == TM_GENERIC)) {
            // a[m] = Z[i1 + m*TPL] of the packed half-length transform
            cx* row = sm + t1 * LP;
            C::sync(grp);
#pragma unroll
            for (int m = 0; m < E; ++m) row[i1 + m * TPL] = a[m];
            C::sync(grp);
            cx* __restrict__ dst = reinterpret_cast<cx*>(p.out.ptr) + off;
            const cx wi = reinterpret_cast<const cx*>(p.rtw)[i1];
            const T h = (T)0.5 * scale;
#pragma unroll
            for (int m = 0; m < E / 2; ++m) {
                const int k = i1 + m * TPL;
                const cx zk = a[m];
                const cx zp = row[(L - k) & (L - 1)];
                const cx A = {zk.x + zp.x, zk.y - zp.y};
                const cx B = {zk.x - zp.x, zk.y + zp.y};
                const cx Cw = cmul(B, cmul(wi, w32<T>(m * (16 / E))));
                const cx xk = {(A.x + Cw.y) * h, (A.y - Cw.x) * h};
                const cx xq = {(A.x - Cw.y) * h, -((A.y + Cw.x) * h)};
                if constexpr (MODE == TM_FAST_R2C) {
                    dst[k] = xk;
                    dst[L - k] = xq;
                } else if (valid) {
                    if ((int64_t)k < p.out.len) dst[(int64_t)k * p.out.elem_stride] = xk;
                    if ((int64_t)(L - k) < p.out.len) dst[(int64_t)(L - k) * p.out.elem_stride] = xq;
                }
            }
            if (i1 == 0) {
                const cx z = a[E / 2];
                const cx v = {z.x * scale, -(z.y * scale)};
                if constexpr (MODE == TM_FAST_R2C) {
                    dst[L / 2] = v;
                } else if (valid && (int64_t)(L / 2) < p.out.len) {
                    dst[(int64_t)(L / 2) * p.out.elem_stride] = v;
                }
            }
        }
        return;
    }

    if (p.st_op == ST_TW) {
        fourstep_twiddle<T, E, TPL>(a, p.tw_lo, p.tw_hi, p.tw_shift, (p.flags & F_TW_CONJ) != 0, i1, lo);
    } else if (p.st_op == ST_MUL && (p.flags & F_CHIRP_GEN)) {
        // cropped positions (>= len) are multiplied by some unit-modulus value and never stored
        chirp_apply<T, E>(a, p, (int64_t)i1 * p.out.pos_es + pos0, (int64_t)TPL * p.out.pos_es, p.chirp_q_out);
    } else if (p.st_op == ST_MUL) {
        const cx* __restrict__ aux = reinterpret_cast<const cx*>(p.aux_out);
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const int64_t pos = (int64_t)(i1 + m * TPL) * p.out.pos_es + pos0;
            if (pos < p.out.len) a[m] = cmul(a[m], aux[pos]);
        }
    }
    if (p.scale != 1.0) {
#pragma unroll
        for (int m = 0; m < E; ++m) a[m] = {a[m].x * scale, a[m].y * scale};
    }
    if (p.flags & F_CONJ_ST_POST) {
#pragma unroll
        for (int m = 0; m < E; ++m) a[m].y = -a[m].y;
    }
    SFC_PHASE_FORCE(a);
    SFC_PHASE(5);  // store operators (twiddle / chirp / scale)
    if constexpr (FAST) {
        if (p.peer_shift >= 0) {
            // fused transpose: each block of the transform axis is stored straight into its
            // destination rank's buffer (peer memory over NVLink, or a local send buffer)
            const int emask = (1 << p.peer_shift) - 1;
#pragma unroll
            for (int m = 0; m < E; ++m) {
                const int e = i1 + m * TPL;
                cx* dst = reinterpret_cast<cx*>(p.peer_out[e >> p.peer_shift]) + off +
                          (int64_t)(e & emask) * p.out.elem_stride;
                *dst = a[m];
            }
            return;
        }
        if constexpr (L <= 64 && MODE == TM_FAST_C2C) {
            if (p.flags & F_STAGE_OUT) {
                constexpr int SP = L + (TPL < 8 ? TPL : 8);
                C::sync(grp);  // the last exchange has been read by everybody
#pragma unroll
                for (int m = 0; m < E; ++m) sm[t1 * SP + i1 + m * TPL] = a[m];
                C::sync(grp);
                cx* __restrict__ tb = reinterpret_cast<cx*>(p.out.ptr) + (int64_t)batch * p.out.batch_stride + (int64_t)tile * (TL * L);
#pragma unroll
                for (int k = 0; k < E; ++k) {
                    const int f = tid + k * C::NT;
                    tb[f] = sm[(f / L) * SP + (f % L)];
                }
                return;
            }
        }
        cx* __restrict__ dst = reinterpret_cast<cx*>(p.out.ptr) + off + (int64_t)i1 * p.out.elem_stride;
        const int64_t step = (int64_t)TPL * p.out.elem_stride;
        if (!(p.flags & F_OUT_NOMASK)) {
            // optional crop: only positions < len are stored
            const int64_t rem = p.out.len - pos0;
            const int elim =
                rem <= 0 ? 0 : (int)min((uint32_t)L, ((uint32_t)rem + (uint32_t)p.out.pos_es - 1u) / (uint32_t)p.out.pos_es);
#pragma unroll
            for (int m = 0; m < E; ++m)
                if (i1 + m * TPL < elim) dst[m * step] = a[m];
        } else {
#pragma unroll
            for (int m = 0; m < E; ++m) dst[m * step] = a[m];
        }
        SFC_PHASE(6);  // store issue
    } else if (p.flags & F_ST_REAL) {
        T* __restrict__ dst = reinterpret_cast<T*>(p.out.ptr) + off;
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const int e = i1 + m * TPL;
            const int64_t pos = (int64_t)e * p.out.pos_es + pos0;
            if (valid && pos < p.out.len) dst[(int64_t)e * p.out.elem_stride] = a[m].x;
        }
    } else {
        cx* __restrict__ dst = reinterpret_cast<cx*>(p.out.ptr) + off;
#pragma unroll
        for (int m = 0; m < E; ++m) {
            const int e = i1 + m * TPL;
            const int64_t pos = (int64_t)e * p.out.pos_es + pos0;
            if (valid && pos < p.out.len) dst[(int64_t)e * p.out.elem_stride] = a[m];
        }
    }
}

template <typename T, int L, int TL, bool DOUBLE, int EMAX = 16, int MODE = TM_GENERIC, int GROUPS = 1>
__global__ void __launch_bounds__(TileCfg<T, L, TL, EMAX, GROUPS>::NT, TileCfg<T, L, TL, EMAX, GROUPS>::MINB)
tile_fft_kernel(const __grid_constant__ PassParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Cx<T>* sm = reinterpret_cast<Cx<T>*>(smem_raw);
    if constexpr (MODE == TM_PIPE_LATE) {
        using C = TileCfg<T, L, TL, EMAX, GROUPS, false>;
        const uint32_t bar = smem_u32(smem_raw + C::SMEM);  // one mbarrier behind the exchange buffer
        if (threadIdx.x == 0) mbar_init(bar, 1);
        __syncthreads();
        if (threadIdx.x < 32 && blockIdx.x < p.total_tiles) pipe_issue<T, L, TL>(p, blockIdx.x, smem_u32(sm), bar);
        uint32_t parity = 0;
        for (uint32_t blk = blockIdx.x; blk < p.total_tiles; blk += gridDim.x, parity ^= 1u)
            tile_body<T, L, TL, DOUBLE, EMAX, MODE, GROUPS>(p, blk, sm, sm, bar, parity, blk + gridDim.x, bar);
    } else if constexpr (MODE != TM_PIPE_C2C) {
        tile_body<T, L, TL, DOUBLE, EMAX, MODE, GROUPS>(p, blockIdx.x, sm, nullptr, 0u, 0u, 0u, 0u);
    } else if constexpr (GROUPS == 1) {
        using C = TileCfg<T, L, TL, EMAX, GROUPS, true>;
        const Cx<T>* land = reinterpret_cast<const Cx<T>*>(smem_raw + C::XCH_BYTES);
        const uint32_t bar = smem_u32(smem_raw + C::XCH_BYTES + C::LAND_BYTES);
        if (threadIdx.x == 0) mbar_init(bar, 1);
        __syncthreads();
        if (threadIdx.x < 32 && blockIdx.x < p.total_tiles) pipe_issue<T, L, TL>(p, blockIdx.x, smem_u32(land), bar);
        uint32_t parity = 0;
        for (uint32_t blk = blockIdx.x; blk < p.total_tiles; blk += gridDim.x, parity ^= 1u)
            tile_body<T, L, TL, DOUBLE, EMAX, MODE, GROUPS>(p, blk, sm, land, bar, parity, blk + gridDim.x, bar);
    } else {
        // Group-pipelined: two thread groups, each transforming its own tile (TLG lanes) out of its own exchange buffer;
        // ONE landing buffer, filled by the TMA unit with the tiles blockIdx.x, +gridDim.x, ... in order.  Tile j is
        // consumed by group j % 2, which then starts the copy of tile j + 1 for the other group: while both groups
        // compute, the next tile is always in flight.
        using C = TileCfg<T, L, TL, EMAX, GROUPS, false>;
        const Cx<T>* land = reinterpret_cast<const Cx<T>*>(smem_raw + C::XCH_BYTES);
        const uint32_t bar0 = smem_u32(smem_raw + C::XCH_BYTES + C::LAND_G_BYTES);
        const int grp = threadIdx.x / C::NTG;
        if (threadIdx.x == 0) {
            mbar_init(bar0, 1);
            mbar_init(bar0 + 8, 1);
        }
        __syncthreads();
        if (threadIdx.x < 32 && blockIdx.x < p.total_tiles) pipe_issue<T, L, C::TLG>(p, blockIdx.x, smem_u32(land), bar0);
        const uint32_t mine = bar0 + 8u * (uint32_t)grp, other = bar0 + 8u * (uint32_t)(grp ^ 1);
        uint32_t k = 0;
        for (uint32_t blk = blockIdx.x + (uint32_t)grp * gridDim.x; blk < p.total_tiles; blk += 2 * gridDim.x, ++k)
            tile_body<T, L, TL, DOUBLE, EMAX, MODE, GROUPS>(p, blk, sm, land, mine, k & 1u, blk + gridDim.x, other);
    }
}

}  // namespace sfc
