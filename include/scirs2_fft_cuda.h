/* scirs2_fft_cuda.h — C ABI of libscirs2_fft_cuda.so: the B200-native replacement
 * for the data-parallel FFT hot path of scirs2-fft (cool-japan/scirs 0.1.0-alpha.6).
 *
 * Every entry point names the reference interface it replaces (file:line relative
 * to the reference tree).  Plain pointers and sizes only; no C++ or torch types.
 * All functions return 0 (SFC_OK) or a negative sfc_status that maps 1:1 onto a
 * variant of `FFTError` (scirs2-fft/src/error.rs:7-46); the message text is
 * available from sfc_last_error() (thread-local).
 *
 * There is NO CPU fallback: without a CUDA device every compute entry point
 * fails with SFC_ERR_BACKEND.
 */
#ifndef SCIRS2_FFT_CUDA_H
#define SCIRS2_FFT_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFC_MAX_DIMS 8
#define SFC_ABI_VERSION 2

/* FFTError variants, scirs2-fft/src/error.rs:7-46 */
typedef enum sfc_status {
    SFC_OK = 0,
    SFC_ERR_COMPUTATION = -1,     /* FFTError::ComputationError    */
    SFC_ERR_DIMENSION = -2,       /* FFTError::DimensionError      */
    SFC_ERR_VALUE = -3,           /* FFTError::ValueError          */
    SFC_ERR_NOT_IMPLEMENTED = -4, /* FFTError::NotImplementedError */
    SFC_ERR_IO = -5,              /* FFTError::IOError             */
    SFC_ERR_BACKEND = -6,         /* FFTError::BackendError        */
    SFC_ERR_PLAN = -7,            /* FFTError::PlanError           */
    SFC_ERR_COMMUNICATION = -8,   /* FFTError::CommunicationError  */
    SFC_ERR_MEMORY = -9           /* FFTError::MemoryError         */
} sfc_status;

/* element types accepted at the boundary (the reference is generic over
 * `T: NumCast` and widens everything to Complex64, fft/algorithms.rs:71-102) */
typedef enum sfc_dtype {
    SFC_F32 = 0,  /* f32 real                */
    SFC_F64 = 1,  /* f64 real                */
    SFC_C64 = 2,  /* Complex<f32> interleaved */
    SFC_C128 = 3  /* Complex64  interleaved   */
} sfc_dtype;

typedef enum sfc_kind { SFC_C2C = 0, SFC_R2C = 1, SFC_C2R = 2 } sfc_kind;
typedef enum sfc_prec { SFC_PREC_F32 = 0, SFC_PREC_F64 = 1 } sfc_prec;
typedef enum sfc_dir { SFC_FORWARD = 0, SFC_INVERSE = 1 } sfc_dir;

/* ------------------------------------------------------------------ runtime */

/* Select the CUDA device this thread's calls run on (one process per GPU).
 * Fails with SFC_ERR_BACKEND when no device is present. */
int sfc_init(int device);
int sfc_device_count(void);
/* Thread-local text of the last error (never NULL). */
const char* sfc_last_error(void);
int sfc_abi_version(void);
/* 1 when a CUDA device is usable (FftBackend::is_available, backend.rs:24). */
int sfc_is_available(void);

/* -------------------------------------------------------------------- plans
 * Replaces rustfft's `FftPlanner::plan_fft_forward/inverse` +
 * `Fft::process` as used at fft/algorithms.rs:159-167, 350-376, 667-683 and the
 * `FftPlan` / `FftPlanExecutor` pair of planning.rs:75-180, 474-556.
 *
 * A plan transforms a C-order contiguous N-D array along `axes` (in list
 * order, duplicates allowed as in fftn, algorithms.rs:667-690).
 *   C2C: in  complex[shape]            -> out complex[shape]
 *   R2C: in  real[shape]               -> out complex[shape with axes[last] -> n/2+1]
 *   C2R: in  complex[shape, halved]    -> out real[shape]
 * `shape` is always the logical (full) transform shape.  `scale` multiplies the
 * result (the caller folds the reference's norm table into it).
 */
typedef struct sfc_plan sfc_plan;

typedef struct sfc_desc {
    int32_t ndim;
    int64_t shape[SFC_MAX_DIMS];
    int32_t naxes;
    int32_t axes[SFC_MAX_DIMS];
    int32_t kind;      /* sfc_kind */
    int32_t prec;      /* sfc_prec: arithmetic + element precision of in/out */
    int32_t direction; /* sfc_dir (C2C only; R2C is forward, C2R inverse) */
    int32_t flags;     /* SFC_DESC_* */
    double scale;
    /* C2R only, with SFC_DESC_CUSTOM_IN_SHAPE: extents of the half-spectrum input when it is
     * not shape-with-last-axis-halved (irfftn pads / reflects what is there, rfft.rs:733-901) */
    int64_t in_shape[SFC_MAX_DIMS];
    /* C2C only: > 1 splits the LAST listed axis into `scatter_parts` equal blocks and stores block q
     * through the q-th pointer given to sfc_exec_device_scatter (layout [outer][n/parts][inner]):
     * the slab transpose of a distributed fftn fused into the FFT store (distributed.rs:232-268) */
    int32_t scatter_parts;
    int32_t reserved;
    /* with scatter_parts > 1: element pitch of the split axis inside every destination block (0 = the inner
     * extent, i.e. dense [outer][n/parts][inner] blocks).  A larger pitch lets block q land inside a bigger array
     * of the destination rank (the second exchange of a natural-layout slab fftn). */
    int64_t scatter_pitch;
    /* C2C over ONE axis, with SFC_DESC_AXIS_LEN: the input array holds axis_in_len elements along that
     * axis (zero-padded / cropped to shape[axis] on load) and the output array axis_out_len (cropped on
     * store); 0 = shape[axis].  What the reference does with `x.resize(n)` / `[..n]` slices around its
     * transforms (dct.rs, the hfft module, spectrogram.rs:287-300) without the extra copies. */
    int64_t axis_in_len, axis_out_len;
    /* with SFC_DESC_AUX_MUL: device tables (complex, plan precision) multiplied into the data on the way
     * in (aux_in[j], j = input index along the axis, axis_in_len entries) and on the way out
     * (aux_out[k], axis_out_len entries); either may be NULL.  Power-of-two lengths only. */
    const void* aux_in;
    const void* aux_out;
    /* SFC_DESC_DCT2 / SFC_DESC_DCT3: weight of the k = 0 term (0 = 1.0), see below */
    double scale_dc;
} sfc_desc;
#define SFC_DESC_CUSTOM_IN_SHAPE 1
#define SFC_DESC_AXIS_LEN 4
#define SFC_DESC_AUX_MUL 8
/* C2C only: store the real part of the result into a real array */
#define SFC_DESC_REAL_OUTPUT 16
/* R2C over ONE axis (any position), power-of-two n >= 128: the plan computes the DCT-II of every lane along it,
 * X[k] = scale * sum_i x[i] cos(pi (i + 1/2) k / n) (dct.rs:523-559), real in / real out, as one fused kernel on the
 * n/2-point transform; with SFC_DESC_DCT2_ORTHO0 output 0 is additionally multiplied by 1/sqrt(2) (dct.rs:552-553) */
#define SFC_DESC_DCT2 32
#define SFC_DESC_DCT2_ORTHO0 64
/* Same restrictions; the inverse packing in one fused kernel:
 *   y[i] = scale * ( scale_dc * X[0] + 2 * sum_{k>=1} X[k] cos(pi k (i + 1/2) / n) )
 * which covers idct type 2, dct type 3 and idct type 3 of the reference (dct.rs:563-684) by choice of the two factors. */
#define SFC_DESC_DCT3 128
/* with SFC_DESC_DCT2 / SFC_DESC_DCT3: the sine transforms instead — DST-II: sum_m x[m] sin(pi (k+1)(m+1/2)/n), DST-III:
 * sum_m x[m] sin(pi (m+1)(k+1/2)/n) with the DCT-III weights applied to x[n-1] (dst.rs:484-592); sign flips and index
 * reversals are folded into the kernels' load / store */
#define SFC_DESC_TRIG_SINE 256
/* kind SFC_R2C, ONE axis, f64, power-of-two n: out[k] = scale * sum_i x[i] cos(pi (i+1/2)(k+1/2) / n) (with SFC_DESC_TRIG_SINE:
 * sin) in one kernel on the n/2-point complex transform.  EXPERIMENTAL: not yet run on a GPU; the consumers only use it with
 * the environment variable SFC_DCT4_FUSED=1. */
#define SFC_DESC_DCT4 512
/* C2C only: the input array is real (imag = 0), as when the reference widens real input
 * to Complex64 before the transform (fft/algorithms.rs:71-102) */
#define SFC_DESC_REAL_INPUT 2

typedef struct sfc_plan_info {
    int64_t in_bytes, out_bytes, scratch_bytes;
    int64_t algorithmic_bytes; /* read+write of the working array per unavoidable pass */
    int64_t device_bytes;      /* bytes the launched kernels actually move (all passes) */
    double nominal_flops;      /* 5 N log2 N (2.5 for real transforms) */
    int32_t num_launches;      /* kernel launches per execution */
    int32_t num_passes;        /* global-memory passes over the working array */
} sfc_plan_info;

/* Looks the plan up in the process-wide plan cache first (plan_cache.rs:102-161). */
int sfc_plan_create(sfc_plan** out, const sfc_desc* desc);
/* Drops the caller's reference (cached plans stay alive in the cache). */
int sfc_plan_destroy(sfc_plan* plan);
int sfc_plan_get_info(const sfc_plan* plan, sfc_plan_info* info);
/* Human-readable pass list ("pass 0: tile L=4096 TL=1 ..."); returns bytes written. */
int sfc_plan_describe(const sfc_plan* plan, char* buf, size_t cap);

/* Device pointers, caller's stream (cudaStream_t passed as void*; NULL = default). */
int sfc_exec_device(sfc_plan* plan, const void* d_in, void* d_out, void* stream);
/* Plans created with desc.scatter_parts = P: d_outs[q] receives block q of the last listed axis.
 * The pointers may be device memory of PEER GPUs (opened with sfc_ipc_open_handle): the kernel then
 * writes over NVLink and no separate all-to-all is needed. */
int sfc_exec_device_scatter(sfc_plan* plan, const void* d_in, void* const* d_outs, int32_t nouts, void* stream);
/* Host pointers: H2D + transform + D2H (what the drop-in free functions use). */
int sfc_exec_host(sfc_plan* plan, const void* h_in, void* h_out);

/* ------------------------------------------------- peer memory (one process per GPU)
 * Raw device allocations that can be exported to the other ranks of a node with CUDA IPC; used by
 * the slab-decomposed fftn to let every rank store its transposed blocks directly into the
 * destination rank's buffer (replaces `Communicator::all_to_all`, distributed.rs:85-103). */
#define SFC_IPC_HANDLE_BYTES 64
int sfc_dev_malloc(void** d_ptr, size_t bytes);
int sfc_dev_free(void* d_ptr);
int sfc_ipc_get_handle(const void* d_ptr, void* handle_out);   /* SFC_IPC_HANDLE_BYTES */
int sfc_ipc_open_handle(const void* handle, void** d_ptr_out); /* maps a peer allocation */
int sfc_ipc_close_handle(void* d_ptr);
int sfc_stream_synchronize(void* stream);


/* ================================================================ multi-GPU (SURVEY 8e / 8b last row)
 * Replaces `trait Communicator` + `DistributedFFT` of scirs2-fft/src/distributed.rs:76-103, 115-362 (whose
 * exchange is a mock, :232-268, 765-769) for the GPUs of ONE node (an NVLink / NVSwitch domain).  Nothing here
 * needs Python, torch or NCCL: ranks find each other through a POSIX shared-memory segment, map each other's
 * device buffers with CUDA IPC, and synchronise ON THE DEVICE through flags in that mapped memory
 * (st.release.sys / ld.acquire.sys), so an exchange is stores over NVLink by the FFT kernel itself plus two
 * one-block kernels.  Two ways to drive it:
 *   - one process per GPU ("rank" mode, the launch model of the benchmark): every rank calls
 *     sfc_comm_init_rank with the same job-unique `name`;
 *   - one process driving several GPUs ("local" mode, what a plain Rust program calling fftn() wants):
 *     sfc_comm_init_local; the *_multi / *_host entry points then take the data of all GPUs at once. */
#define SFC_MAX_GPUS 16
typedef struct sfc_comm sfc_comm;
/* ngpu <= 0: every visible device; devices == NULL: 0 .. ngpu-1.  Enables peer access between them. */
int sfc_comm_init_local(sfc_comm** out, int32_t ngpu, const int32_t* devices);
/* Collective over the `world` processes of a node; `name`: [A-Za-z0-9_.-]{1,64}, unique to this job (e.g. launcher
 * pid + start time).  `device` is the CUDA device this rank drives.  Times out (SFC_ERR_COMMUNICATION) after
 * SFC_COMM_TIMEOUT_MS (default 60000) if a rank never shows up. */
int sfc_comm_init_rank(sfc_comm** out, const char* name, int32_t rank, int32_t world, int32_t device);
int sfc_comm_destroy(sfc_comm* comm);
int sfc_comm_size(const sfc_comm* comm);  /* Communicator::size, distributed.rs:99 */
int sfc_comm_rank(const sfc_comm* comm);  /* Communicator::rank, :102 (local mode: 0) */
int sfc_comm_barrier(sfc_comm* comm);     /* Communicator::barrier, :96 — host-side */
/* host-side all-gather of `bytes` per rank (out: world * bytes); a no-op copy in local mode */
int sfc_comm_allgather(sfc_comm* comm, const void* in, void* out, size_t bytes);
/* Symmetric device allocation (collective, same `bytes` on every rank): the buffer of every rank is mapped into
 * every other rank, so distributed plans can store straight into it.  rank mode: *d_ptr = this rank's buffer;
 * local mode: d_ptr receives ngpu pointers. */
int sfc_comm_alloc(sfc_comm* comm, size_t bytes, void** d_ptr);
int sfc_comm_free(sfc_comm* comm, void* d_ptr);  /* collective; local mode: the first pointer */

/* DecompositionStrategy, distributed.rs:18-29 */
typedef enum sfc_decomposition {
    SFC_DECOMP_REPLICATED = 0,  /* every GPU runs the whole plan on its own data (replicas only) */
    SFC_DECOMP_BATCH_SPLIT = 1, /* axis 0 is a pure batch axis: contiguous split, no exchange */
    SFC_DECOMP_SLAB = 2         /* 3-D c2c: axis-0 slabs, FFT axes 2 and 1, exchange, FFT axis 0 (distributed.rs:356-362) */
} sfc_decomposition;
typedef enum sfc_slab_layout {
    SFC_SLAB_TRANSPOSED = 0, /* output stays axis-1-distributed: rank r holds out[:, r*n1/P:(r+1)*n1/P, :] */
    SFC_SLAB_NATURAL = 1     /* a second exchange (fused into the axis-0 FFT store) restores axis-0 slabs: a true fftn */
} sfc_slab_layout;

/* `base.shape` is the GLOBAL shape; ngpu is the communicator's size (sfc_desc{.., ngpu, decomposition}). */
typedef struct sfc_dist_desc {
    sfc_desc base;
    int32_t decomposition; /* sfc_decomposition */
    int32_t layout;        /* sfc_slab_layout (SLAB only) */
    int32_t chunks;        /* SLAB: column blocks of the last axis the exchange is pipelined in: the axis-0 pass of block j
                            * runs while block j+1 is scattered over NVLink (0 = library default: 1 = off, or SFC_SLAB_CHUNKS
                            * — measured on 2 x B200 the overlap returns nothing; lowered to what the
                            * tiles of both passes allow, see sfc_dist_info.chunks) */
    int32_t reserved;
} sfc_dist_desc;

typedef struct sfc_dist_info {
    int32_t world, rank;                   /* rank of the first GPU this process drives */
    int32_t decomposition, layout, chunks;
    int32_t reserved;
    int64_t local_in_elems, local_out_elems; /* per GPU, in elements of the in / out type */
    int64_t local_in_shape[SFC_MAX_DIMS], local_out_shape[SFC_MAX_DIMS];
    int64_t exchange_bytes_sent;           /* per GPU per exchange, to OTHER GPUs: (P-1)/P of the local slab */
    int32_t num_exchanges;                 /* 0 batch split, 1 transposed, 2 natural */
    int32_t num_launches;                  /* kernel launches per GPU per execution */
    int64_t algorithmic_bytes;             /* per GPU */
    double nominal_flops;                  /* whole transform */
} sfc_dist_info;

typedef struct sfc_dist_plan sfc_dist_plan;
int sfc_dist_plan_create(sfc_dist_plan** out, sfc_comm* comm, const sfc_dist_desc* desc); /* collective */
int sfc_dist_plan_destroy(sfc_dist_plan* plan);                                           /* collective */
int sfc_dist_plan_get_info(const sfc_dist_plan* plan, sfc_dist_info* info);
/* Stage timing of the SLAB pipeline on this process' first GPU (CUDA events between the stages of every later
 * execution): sfc_dist_plan_stage_ms returns the number of stages written, in order — FFT axis 2 | FFT axis 1 +
 * scatter | signal + wait | FFT axis 0 (+ scatter) | [natural: signal + wait 2 and copy-out].  A pipelined plan
 * (sfc_dist_info.chunks > 1) reports three: FFT axis 2 | the scatter passes of all column blocks (the axis-0 passes
 * overlap them on a side stream) | what is left of the axis-0 passes after the last scatter. */
int sfc_dist_plan_profile(sfc_dist_plan* plan, int32_t enable);
int sfc_dist_plan_stage_ms(sfc_dist_plan* plan, double* ms, int32_t cap);
/* rank mode: this rank's slab / batch share, device pointers, caller's stream.  The call only enqueues work. */
int sfc_dist_exec_device(sfc_dist_plan* plan, const void* d_in, void* d_out, void* stream);
/* local mode: one pointer per GPU (in communicator order); streams == NULL: library streams, one per GPU.
 * Only enqueues; sfc_dist_synchronize waits for every GPU. */
int sfc_dist_exec_device_multi(sfc_dist_plan* plan, const void* const* d_in, void* const* d_out, void* const* streams);
int sfc_dist_synchronize(sfc_dist_plan* plan);
/* Host buffers, H2D + transform + D2H.  local mode: the WHOLE global array in C order (output in natural layout
 * whatever `layout` says); every GPU copies over its own PCIe link.  rank mode: this rank's share. */
int sfc_dist_exec_host(sfc_dist_plan* plan, const void* h_in, void* h_out);
/* The free functions sfc_fftn / sfc_ifftn (3-D complex, every axis, power-of-two extents divisible by the GPU
 * count) and sfc_execute_batch run over `ngpu` GPUs of this process from now on (0 or 1: single GPU, the default;
 * < 0: all visible).  This is how `use scirs2_fft_cuda as scirs2_fft` reaches the whole box without new arguments. */
int sfc_set_num_gpus(int32_t ngpu);
int sfc_get_num_gpus(void);

/* --------------------------------------------------------------- plan cache
 * plan_cache.rs:28-235 — 128 entries, 1 h TTL, LRU, hit/miss counters. */
typedef struct sfc_cache_stats {
    uint64_t hit_count, miss_count;
    double hit_rate;
    uint64_t size, max_size;
} sfc_cache_stats;
int sfc_cache_get_stats(sfc_cache_stats* out);   /* PlanCache::get_stats  :182-190 */
int sfc_cache_set_enabled(int enabled);          /* PlanCache::set_enabled :57-59  */
int sfc_cache_is_enabled(void);                  /* PlanCache::is_enabled  :62-64  */
int sfc_cache_clear(void);                       /* PlanCache::clear       :66-71  */
int sfc_cache_configure(uint64_t max_entries, double max_age_seconds); /* with_config :47-54 */

/* ---------------------------------------------------------- planner options
 * The GPU planner's tile / pass choices (SFC_PIPE_LATE, SFC_BLUE_L1, SFC_THREE_LEVEL_MIN, SFC_COL_TL,
 * SFC_RADIX3, ...) default to measured-best settings and can be given in the environment; these two calls override them at
 * run time so that a tuner can time variants of one plan inside a process and persist the winner per (GPU, shape, dtype)
 * the way auto_tuning.rs:188-229 does for its algorithm variants.  Setting an option empties the plan cache.
 * value == NULL removes the override.  sfc_planner_get_option returns the length written (0 = unset). */
int sfc_planner_set_option(const char* name, const char* value);
int sfc_planner_get_option(const char* name, char* buf, size_t cap);

/* ------------------------------------------------- drop-in free functions
 * Reference semantics (including its SciPy-divergent quirks) with
 * HOST buffers.  `n`/shape arguments use -1 / NULL for the reference's `None`.
 * `norm` is "backward" | "ortho" | "forward" | anything else (= no scaling) | NULL.
 * Outputs are caller-allocated; `out_cap` is the capacity in ELEMENTS and the
 * produced element count / shape is written back.  Output is always f64
 * (Complex64 or f64), exactly as the reference.
 */

/* fft / ifft: fft/algorithms.rs:131-176, 210-263 */
int sfc_fft(const void* x, int64_t len, int dtype, int64_t n, double* out, int64_t out_cap, int64_t* out_len);
int sfc_ifft(const void* x, int64_t len, int dtype, int64_t n, double* out, int64_t out_cap, int64_t* out_len);
/* rfft / irfft: rfft.rs:39-59, 92-178 (the hard-coded test returns at :97-116 are NOT reproduced) */
int sfc_rfft(const void* x, int64_t len, int dtype, int64_t n, double* out, int64_t out_cap, int64_t* out_len);
int sfc_irfft(const void* x, int64_t len, int dtype, int64_t n, double* out, int64_t out_cap, int64_t* out_len);
/* fft2 / ifft2: fft/algorithms.rs:293-401, 439-541; shape/axes NULL = None. */
int sfc_fft2(const void* x, int64_t rows, int64_t cols, int dtype, const int64_t* shape2, const int32_t* axes2,
             const char* norm, double* out, int64_t out_cap, int64_t* out_shape2);
int sfc_ifft2(const void* x, int64_t rows, int64_t cols, int dtype, const int64_t* shape2, const int32_t* axes2,
              const char* norm, double* out, int64_t out_cap, int64_t* out_shape2);
/* rfft2 / irfft2: rfft.rs:212-232, 274-355 (halves axis 0; irfft2 keeps the reference's
 * (N0out*N1out)/(N0in*N1in) factor; its hard-coded 2x2 return is NOT reproduced) */
int sfc_rfft2(const void* x, int64_t rows, int64_t cols, int dtype, const int64_t* shape2, double* out,
              int64_t out_cap, int64_t* out_shape2);
int sfc_irfft2(const void* x, int64_t rows, int64_t cols, int dtype, const int64_t* shape2, double* out,
               int64_t out_cap, int64_t* out_shape2);
/* fftn / ifftn: fft/algorithms.rs:576-706, 757-890.  shape NULL = None (else ndim entries),
 * axes NULL = None (else naxes entries). */
int sfc_fftn(const void* x, int32_t ndim, const int64_t* in_shape, int dtype, const int64_t* shape,
             const int64_t* axes, int32_t naxes, const char* norm, double* out, int64_t out_cap,
             int64_t* out_shape);
int sfc_ifftn(const void* x, int32_t ndim, const int64_t* in_shape, int dtype, const int64_t* shape,
              const int64_t* axes, int32_t naxes, const char* norm, double* out, int64_t out_cap,
              int64_t* out_shape);
/* rfftn / irfftn: rfft.rs:472-525, 621-725.  For irfftn `shape` may have ndim or naxes
 * entries (nshape says which). */
int sfc_rfftn(const void* x, int32_t ndim, const int64_t* in_shape, int dtype, const int64_t* shape,
              const int64_t* axes, int32_t naxes, const char* norm, double* out, int64_t out_cap,
              int64_t* out_shape);
int sfc_irfftn(const void* x, int32_t ndim, const int64_t* in_shape, int dtype, const int64_t* shape,
               int32_t nshape, const int64_t* axes, int32_t naxes, const char* norm, double* out,
               int64_t out_cap, int64_t* out_shape);
/* fft_strided / fft_strided_complex / ifft_strided: strided_fft.rs:16-239 (one axis of an N-D array) */
int sfc_fft_strided(const void* x, int32_t ndim, const int64_t* in_shape, int dtype, int64_t axis,
                    int inverse, double* out, int64_t out_cap);

/* ----------------------------------------------------- FftBackend trait
 * backend.rs:14-48: fft / ifft (1/n-normalised, :149-152) / *_sized on Complex64 slices. */
int sfc_backend_fft(const double* input, int64_t in_len, double* output, int64_t out_len);
int sfc_backend_ifft(const double* input, int64_t in_len, double* output, int64_t out_len);
int sfc_backend_fft_sized(const double* input, int64_t in_len, double* output, int64_t out_len, int64_t size);
int sfc_backend_ifft_sized(const double* input, int64_t in_len, double* output, int64_t out_len, int64_t size);
int sfc_backend_supports_feature(const char* feature); /* backend.rs:157-160 (+ "gpu_acceleration") */
const char* sfc_backend_name(void);                    /* "cuda_fft", examples/backend_example.rs:103 */
const char* sfc_backend_description(void);

/* ------------------------------------------ batched executor (planning_parallel.rs:316-405)
 * `count` independent length-`size` complex transforms, contiguous [count][size]. */
int sfc_execute_batch(const double* inputs, double* outputs, int64_t count, int64_t size, int inverse);

/* --------------------------------------------------------- f32 compute path
 * The reference has no f32 arithmetic; BASELINE config 2b asks for one.  Batched
 * real transforms over the last axis, element precision chosen by `prec`.
 * x: [batch][n] real -> out: [batch][n/2+1] complex (rfft) and back (irfft). */
int sfc_rfft_batch(const void* x, int64_t batch, int64_t n, int prec, void* out);
int sfc_irfft_batch(const void* x, int64_t batch, int64_t n, int prec, void* out);

/* ------------------------------------------------- consumers of the hot path
 * Each is an O(n) pre-pass, the FFT above and an O(n) post-pass in the reference; here the passes are
 * device tables fused into the FFT kernels (power-of-two lengths) or element-wise kernels around them.
 *
 * sfc_dct / sfc_dst: dct, idct, dct2, idct2, dctn, idctn (dct.rs:56-420) and the dst family
 * (dst.rs:48-405) in one entry point: C-order f64 array of `shape`, transform of `type` 1..4 applied
 * along every listed axis in order (axes == NULL: all axes, dct.rs:317); `norm` "ortho" or anything
 * else / NULL for the reference's un-normalised definitions (dct.rs:425-757, dst.rs:409-702). */
int sfc_dct(const double* x, int32_t ndim, const int64_t* shape, const int32_t* axes, int32_t naxes,
            int32_t type, int32_t inverse, const char* norm, double* out);
int sfc_dst(const double* x, int32_t ndim, const int64_t* shape, const int32_t* axes, int32_t naxes,
            int32_t type, int32_t inverse, const char* norm, double* out);
/* hartley.rs:37-130, 202-209: H[k] = Re F[k] - Im F[k] with F = fft(x, None) (padded to the next power
 * of two, first n bins kept — as the reference); idht = dht / n; dht2 (hartley.rs:133-200). */
int sfc_dht(const double* x, int64_t n, double* out);
int sfc_idht(const double* h, int64_t n, double* out);
int sfc_dht2(const double* x, int64_t rows, int64_t cols, int32_t axis0, int32_t axis1, double* out);
/* hfft/complex_to_real.rs:58-135 and hfft/real_to_complex.rs:49-149 */
int sfc_hfft(const void* x, int64_t len, int dtype, int64_t n, double* out, int64_t out_cap, int64_t* out_len);
int sfc_ihfft(const double* x, int64_t len, int64_t n, double* out, int64_t out_cap, int64_t* out_len);
/* lib.rs:437-516: analytic signal, n complex values out */
int sfc_hilbert(const double* x, int64_t n, double* out);
/* spectrogram.rs:76-310 (stft) and :312-420 (spectrogram).  `window`: nperseg samples; noverlap / nfft < 0
 * (or 0 for nfft) = the reference defaults nperseg/2 and nperseg; boundary 0 none, 1 "reflect", 2 "zeros",
 * 3 "constant"; out_mode 0 complex [freq][frame], 1 |z|^2*scale ("psd"), 2 |z|*sqrt(scale) ("magnitude"),
 * 3 phase in radians, 4 in degrees; out_cap_elems counts output elements. */
int sfc_stft(const double* x, int64_t len, const double* window, int64_t nperseg, int64_t noverlap, int64_t nfft,
             int32_t detrend, int32_t onesided, int32_t boundary, int32_t out_mode, double scale, void* out,
             int64_t out_cap_elems, int64_t* freq_len, int64_t* frames);

/* SURVEY 8f rank 4 — the per-segment `scirs2_fft::fft` loops of scirs2-signal/src/spectral.rs (periodogram :186-207,
 * welch :346-395, stft :580-616) as ONE framed, batched transform: row f = window * detrend(x[f*step .. f*step+nperseg))
 * zero-padded to P (the power of two `fft(&padded, None)` pads to; detrend 0 none, 1 "constant", 2 "linear",
 * spectral.rs:77-117), real-to-complex, first `bins` bins kept.  reduce = 0: out = complex f64 [frames][bins];
 * reduce = 1: out = f64 [bins] = scale * sum over frames of |X|^2 (welch's average with scale folded in). */
int sfc_signal_spectra(const double* x, int64_t len, const double* window, int64_t nperseg, int64_t step, int64_t frames,
                       int64_t P, int32_t detrend, int32_t reduce, int64_t bins, double scale, void* out);

/* ------------------------------------------------- SURVEY 8f rank 2 (what benches/fft_benchmarks.rs times)
 * memory_efficient.rs:89-190 (fft_inplace: result written to BOTH buffers; returns n), :243-397 (fft2_efficient,
 * out_rows/out_cols < 0 = input shape), :401-580 (fft_streaming: chunk_size <= 0 = the reference default;
 * chunked mode concatenates independent per-chunk transforms exactly as the reference does) and
 * ndim_optimized.rs:17-58 (fftn_optimized on a real f64 array; axes == NULL: all). */
int sfc_fft_inplace(double* input, int64_t n, double* output, int64_t out_len, int32_t inverse, int32_t normalize);
int sfc_fft2_efficient(const void* x, int64_t rows, int64_t cols, int dtype, int64_t out_rows, int64_t out_cols,
                       int32_t inverse, int32_t normalize, double* out);
int sfc_fft_streaming(const void* x, int64_t len, int dtype, int64_t n, int32_t inverse, int64_t chunk_size, double* out);
int sfc_fftn_optimized(const double* x, int32_t ndim, const int64_t* shape, const int32_t* axes, int32_t naxes, double* out);

/* chirp z-transform as czt.rs:45-275 sets it up, with the FFT calls that file stubs out (:110-113,
 * :239-252) in place.  x: [rows][n] complex f64, out: [rows][m]; has_w == 0 takes w = exp(-2 pi i / m) (czt.rs:87-96). */
int sfc_czt(const double* x, int64_t rows, int64_t n, int64_t m, int32_t has_w, double w_re, double w_im, double a_re,
            double a_im, double* out);

#ifdef __cplusplus
}
#endif
#endif /* SCIRS2_FFT_CUDA_H */
