// plan.h — GPU planner: picks the pass decomposition per transform length and
// owns the device tables / scratch of one plan.
//
// Replaces `FftPlanner` + `Arc<dyn Fft<f64>>` (rustfft, external) as consumed at
// scirs2-fft/src/fft/algorithms.rs:159-167 and the plan objects of
// scirs2-fft/src/planning.rs:75-180.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/scirs2_fft_cuda.h"
#include "aux_kernels.cuh"
#include "kernel_registry.h"
#include "pass_params.h"

namespace sfc {

enum BufRole : int { R_IN = 0, R_OUT = 1, R_SA = 2, R_MS = 3 };
enum StepKind : int { K_TILE = 0, K_COPY = 1, K_HERM = 2 };

// one kernel launch (or one launch per batch chunk)
struct Step {
    int kind = K_TILE;
    const KernelEntry* k = nullptr;
    PassParams p{};
    CopyParams cp{};
    HermParams hp{};
    int src = R_IN, dst = R_OUT;
    size_t src_esize = 16, dst_esize = 16;  // bytes per addressed element (for batch offsets)
    int group = -1;                          // steps sharing a group are chunk-looped together
    bool batch_fastest = false;              // CTA order: batch index fastest (table reuse in L2)
    bool scatter = false;                    // store through the caller's per-block pointer table
    int tmap = 0;                            // late-prefetch column tiles: 1 = lanes contiguous over the batch, 2 = per outer group
    int64_t nbatch = 1;                      // batches (blockIdx-level outer index)
    int64_t batch_mult = 1;                  // grouped steps: blockIdx-level batches per group batch (three-level inner passes)
    std::string desc;
};

struct Group {
    int64_t nbatch = 1;
    int64_t chunk = 1;  // batches per launch round
    // L2 blocking: `ways` rounds are in flight at once, each on its own side stream with its own
    // slice of the work area, small enough that the passes of one round hand their data over in L2
    int ways = 1;
    size_t slice_bytes = 0;  // work-area bytes per round
};

// Execute only column block `index` of `count` of every row of `row_lanes` adjacent lanes (PassParams::win_*).
struct ExecWindow {
    int64_t row_lanes;
    int index, count;
};

struct PlanError {
    int code;
    std::string msg;
};

class Plan {
   public:
    static std::shared_ptr<Plan> create(const sfc_desc& d, PlanError& err);
    ~Plan();

    // d_in / d_out are device pointers.  Safe to call from several threads and on several streams at once: launches are
    // enqueued under the plan's mutex, and an execution on another stream than the previous one first waits (on the
    // device, through an event) for the previous one, because both use the plan's scratch areas.
    int exec(const void* d_in, void* d_out, cudaStream_t stream, std::string& err, void* const* scatter = nullptr,
             int nscatter = 0, const struct ExecWindow* win = nullptr);
    // true when exec() accepts a window of `nwin` column blocks over rows of `row_lanes` lanes (one tile-kernel launch,
    // no scratch, no batch loop, tiles that divide the blocks)
    bool window_ok(int64_t row_lanes, int nwin) const;

    sfc_desc desc{};
    sfc_plan_info info{};
    int device = 0;
    std::string describe() const;

    int64_t in_elems = 0, out_elems = 0;  // elements of the in / out arrays
    size_t in_esize = 16, out_esize = 16;

   private:
    Plan() = default;
    std::vector<Step> steps_;
    std::vector<Group> groups_;
    void* sa_ = nullptr;  // array-layout scratch
    size_t sa_bytes_ = 0;
    void* ms_ = nullptr;  // four-step / Bluestein work area
    size_t ms_bytes_ = 0;
    std::mutex mu_;
    // side streams for L2-blocked rounds (created on first use, on the plan's device)
    std::vector<cudaStream_t> side_;
    std::vector<cudaEvent_t> side_done_;
    cudaEvent_t fork_ev_ = nullptr;
    bool ensure_side_streams(int n, std::string& es);
    // scratch is allocated on first execution (not at plan creation) and can be given back under memory pressure
    bool ensure_scratch(std::string& es);
    cudaEvent_t busy_ev_ = nullptr;      // recorded after the last launch of every execution that touches scratch
    bool busy_valid_ = false;
    cudaStream_t last_stream_ = nullptr;
    friend struct PlanBuilder;

   public:
    // Frees the scratch areas if the plan is idle (no execution being enqueued; waits for the last one to finish on the
    // device).  Returns the bytes released.  Called for the OTHER cached plans when a device allocation fails.
    size_t release_scratch();
    // Allocates the scratch areas now (what the first execution would do): needed before stream capture, and by callers
    // that must not hit a device allocation in the middle of a multi-GPU enqueue.
    int prepare(std::string& es);
    size_t scratch_resident() const { return (sa_ ? sa_bytes_ : 0) + (ms_ ? ms_bytes_ : 0); }
};

// Device-memory pressure: api.cu installs a hook that walks the plan cache and calls release_scratch() on every plan but
// `except`; plan.cu and ensure_buf call alloc_with_relief, which retries a failed cudaMalloc once after running it.
void set_scratch_pressure_hook(size_t (*hook)(const Plan* except));
cudaError_t alloc_with_relief(void** p, size_t bytes, const Plan* except);

// largest single-tile transform per precision
inline int lmax_for(int prec) { return prec == PREC_F64 ? 8192 : 16384; }

// device tables (built once per device, never freed before process exit)
const void* table_stage_tw(int prec, int L, PlanError& err);                // W_L^j, j < L
const void* table_rtw(int prec, int L, PlanError& err);                     // W_{2L}^i, i < max(L/16,1)
bool table_fourstep(int prec, int64_t M, const void** lo, const void** hi, int* shift, PlanError& err);
bool table_chirp_roots(int64_t N, const void** lo, const void** hi, int* shift, PlanError& err);  // R(k) = exp(-i*pi*k/N)
const void* table_chirp(int prec, int64_t N, PlanError& err);               // exp(-i*pi*n^2/N), n < N
// FFT_M(conj chirp, wrapped)/M ; layout: natural (L1 == 0), [k1][k2] four-step order, or [k1][k2][k3] (L3 > 0) three-level order
const void* table_bluestein_b(int prec, int64_t N, int64_t M, int64_t L1, int64_t L2, PlanError& err, int64_t L3 = 0);

void set_error(int code, const std::string& msg);

// planner options (the SFC_* knobs of DESIGN.md section 11): a run-time override takes precedence over the environment
void planner_set_option(const char* name, const char* value);  // value == nullptr: back to the environment / default
std::string planner_get_option(const char* name);

}  // namespace sfc
