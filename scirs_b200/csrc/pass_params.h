// pass_params.h — host/device shared description of ONE global-memory pass.
//
// A "pass" is one launch of the tile FFT kernel (fft_tile.cuh): every CTA loads
// a tile of TL lanes x L elements, transforms each lane in shared memory /
// registers and stores it.  All the reference's wrapper work around the rustfft
// call (gather -> process -> scatter -> scale, scirs2-fft/src/fft/algorithms.rs:
// 677-703) is folded into the load / store operators below.
#pragma once
#include <stdint.h>

namespace sfc {

// thread -> (lane, butterfly) mapping of a tile
enum MapMode : int32_t {
    MAP_ROW = 0,  // butterfly index fastest: lanes are contiguous rows (elem stride 1)
    MAP_COL = 1,  // lane index fastest: adjacent lanes are contiguous in memory
};

enum LoadOp : int32_t {
    LD_C = 0,       // complex element
    LD_R = 1,       // real element, imag = 0            (reference: convert_to_complex, algorithms.rs:71-94)
    LD_C_MUL = 2,   // complex element * aux_in[pos]     (Bluestein chirp pre-multiply)
    LD_R_MUL = 3,   // real element * aux_in[pos]
    LD_C2R = 4,     // Hermitian half-spectrum -> packed N/2 complex (irfft fast path, rfft.rs:92-178)
    LD_SPLIT2 = 5,  // radix-2 DIF pre-stage: lane parity 0 gets x[j] + x[j+L], parity 1 gets (x[j] - x[j+L]) W_2L^j
};

enum StoreOp : int32_t {
    ST_C = 0,       // complex * scale
    ST_TW = 1,      // complex * W_M^(e*lane_outer) * scale   (four-step inter-pass twiddle)
    ST_MUL = 2,     // complex * aux_out[pos] * scale         (Bluestein chirp post-multiply)
    ST_R2C = 3,     // packed N/2 complex spectrum -> N/2+1 Hermitian half (rfft fast path, rfft.rs:39-59)
};

enum PassFlags : uint32_t {
    // inverse transforms run the forward code on conjugated data: IFFT(x) = conj(FFT(conj(x)))
    F_CONJ_LD_PRE = 1u << 0,   // conjugate right after the raw load         (outer inverse)
    F_CONJ_LD_POST = 1u << 1,  // conjugate after the load operator          (inner inverse FFT)
    F_CONJ_ST_PRE = 1u << 2,   // conjugate before the store operator        (undo inner inverse)
    F_CONJ_ST_POST = 1u << 3,  // conjugate just before the raw store        (outer inverse)
    F_TW_CONJ = 1u << 4,       // ST_TW uses conj(W)
    F_ST_REAL = 1u << 5,
    F_IN_NOMASK = 1u << 6,     // every (lane, e) position is < in.len: loads need no bounds predicate
    F_OUT_NOMASK = 1u << 7,
    F_CHIRP_GEN = 1u << 8,
    // fused DCT kernels computing the sine transforms: DST-II(x)[k] = DCT-II((-1)^m x[m])[n-1-k],
    // DST-III(x)[k] = (-1)^k DCT-III(x reversed)[k]   (sign flips and index reversals folded into load / store)
    F_TRIG_SINE = 1u << 9,
    // short contiguous rows (L <= 64): the tile is moved between global and shared memory with fully coalesced
    // 128-bit accesses (thread t takes flat element t, t + NT, ...) and the per-thread rows are read from there;
    // without it thread t walks its own 16..64-element row and every warp request touches 32 cache lines
    F_STAGE_IN = 1u << 10,
    F_STAGE_OUT = 1u << 11,
    // inter-pass twiddle applied on the way IN (W^(e * lane_outer) from ld_tw_*; conjugated with F_LD_TW_CONJ):
    // the inverse half of the three-level Bluestein, whose twiddle index is (element, lane) of the NEXT pass
    F_LD_TW = 1u << 12,
    F_LD_TW_CONJ = 1u << 13,
    // late-prefetch flavour, strided (column) tiles: the tile is landed by ONE-to-FEW cp.async.bulk.tensor copies through
    // the tensor map tmap_in (a 4-D view [batch][outer][element][lane] of the input array) instead of per-row bulk copies
    F_TMAP_IN = 1u << 14,     // LD_*_MUL / ST_MUL: generate the Bluestein chirp in registers instead of reading aux_*    // likewise for stores       // store only the real part into a real array (irfftn, rfft.rs:722)
};

struct IoDesc {
    void* ptr;
    int64_t batch_stride;  // elements, per batch index (blockIdx / tiles_per_batch)
    int64_t outer_stride;  // elements, per (lane / inner_count)
    int64_t inner_stride;  // elements, per (lane % inner_count)
    int64_t elem_stride;   // elements, per transform index e
    int64_t len;           // valid logical positions: pos < len is loaded / stored, else 0 / skipped
    int64_t pos_es;        // logical position pos = e*pos_es + lane_outer*pos_ls  (lane_outer = lane / inner_count)
    int64_t pos_ls;
};

struct PassParams {
    IoDesc in, out;
    uint32_t nlanes;           // lanes per batch
    uint32_t inner_count;      // lanes per outer index (shared by in/out)
    uint32_t tiles_per_batch;  // ceil(nlanes / TL)
    uint32_t nbatch_fast;      // != 0: blockIdx = tile * nbatch_fast + batch (set per launch), else batch * tiles + tile
    uint32_t total_tiles;      // pipelined flavour: tiles of this launch (the grid is smaller and persistent)
    uint32_t tile_group_shift; // with nbatch_fast: 2^shift adjacent tiles stay adjacent in CTA order (DRAM row locality)
    int32_t map_in, map_out;
    int32_t ld_op, st_op;
    uint32_t flags;
    const void* tw;            // W_L^j, j < L              (stage twiddles)
    const void* aux_in;        // LD_*_MUL table, indexed by in pos
    const void* aux_out;       // ST_MUL table, indexed by out pos
    const void* tw_lo;         // ST_TW: W_M^j, j < 2^tw_shift
    const void* tw_hi;         // ST_TW: W_M^(j << tw_shift)
    int32_t tw_shift;
    const void* mid;           // double kernels: pointwise table between the two transforms
    int64_t mid_es, mid_ls;    // mid index = e*mid_es + lane_outer*mid_ls + lane_inner*mid_is
    int64_t mid_is;
    const void* ld_tw_lo;      // F_LD_TW tables (same two-level layout as tw_lo / tw_hi)
    const void* ld_tw_hi;
    int32_t ld_tw_shift;
    const void* rtw;           // R2C/C2R: W_{2L}^i, i < L/E
    // F_CHIRP_GEN: chirp[n] = exp(-i*pi*n^2/N) = R(n^2 mod 2N), R(k) = chirp_hi[k >> shift] * chirp_lo[k & mask]
    // (f64 tables whatever the transform precision); q_* = exp(-i*pi*2*D^2/N) for the thread's position step D
    const void* chirp_lo;
    const void* chirp_hi;
    int32_t chirp_shift;
    uint64_t chirp_mod;        // 2N
    double chirp_q_in[2], chirp_q_out[2];
    double scale;
    double scale_dc;           // TM_FAST_DCT2: extra factor of output 0; TM_FAST_DCT3: factor of input 0
    // split-axis scatter (slab transpose fused into the store): output element e of the transform
    // axis goes to peer_out[e >> peer_shift] at element index (e & ((1 << peer_shift) - 1)).
    // peer_shift < 0 = off.  The pointers may be peer-GPU memory mapped over NVLink.
    int32_t peer_shift;
    void* peer_out[16];
    // F_TMAP_IN: CUtensorMap (opaque, 128 bytes, 64-byte aligned) encoded by Plan::exec for the input pointer of this launch;
    // dims (fastest first) = [2 * lanes-contiguous][L elements][outer][batch] in units of the real type
    alignas(64) unsigned char tmap_in[128];
    int32_t tmap_box_rows;   // elements of the transform axis per box (min(L, 256)); L / tmap_box_rows copies per tile
    int32_t tmap_split;      // 0: lanes are contiguous over the whole batch (coordinate 0 = first lane of the tile);
                             // 1: coordinate 0 = lane inside its outer group, coordinate 2 = the outer index
    // Tile window (single-launch plans, set per execution by Plan::exec): the lanes form rows of win_row_tiles tiles and
    // this launch only takes tiles [win_first, win_first + win_len) of every row — CTA blk works on tile
    // (blk / win_len) * win_row_tiles + win_first + blk % win_len.  win_len == 0 = off.  Lets the pipelined slab exchange
    // run the strided passes chunk by chunk over column blocks of the SAME dense arrays (dist.cu).
    uint32_t win_row_tiles, win_first, win_len;
#ifdef SFC_PHASE_TIMING
    unsigned long long* dbg;   // developer build only: per-launch phase clock sums (thread 0 of every CTA)
#endif
};

}  // namespace sfc
