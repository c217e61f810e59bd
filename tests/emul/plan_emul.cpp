// TEST INFRASTRUCTURE ONLY — the planner (plan.cu) and the f64 kernel translation units compiled for the host with
// -DSFC_HOST_EMUL against tests/emul/cuda_runtime.h, plus this driver: create a plan from an sfc_desc, execute it on host
// arrays, return its description.  Checks planner plumbing and kernel index logic without a GPU (no f32, no TMA-pipelined
// or two-group flavours: those need PTX that is not emulated).
#include <cuda_runtime.h>
#include <memory>
#include <string>

#include "plan.h"
#include "aux_kernels.cuh"

thread_local emul_dim3 threadIdx;
thread_local emul_dim3 blockIdx;
emul_dim3 blockDim, gridDim;
pthread_barrier_t emul_cta_barrier;
namespace sfc {
alignas(128) unsigned char smem_raw[232448];  // the kernel's `extern __shared__` array: one CTA runs at a time

void set_error(int, const std::string&) {}
// flavours that are not emulated register nothing
void register_kernels_f32_small(void (*)(const KernelEntry&)) {}
void register_kernels_f32_mid(void (*)(const KernelEntry&)) {}
void register_kernels_f32_big(void (*)(const KernelEntry&)) {}
void register_kernels_f32_dbl_a(void (*)(const KernelEntry&)) {}
void register_kernels_f32_dbl_b(void (*)(const KernelEntry&)) {}
void register_kernels_f32_real(void (*)(const KernelEntry&)) {}
void register_kernels_e8(void (*)(const KernelEntry&)) {}
void register_kernels_pipe(void (*)(const KernelEntry&)) {}
void register_kernels_pipe_dbl(void (*)(const KernelEntry&)) {}
}  // namespace sfc

extern "C" int emul_plan_run(const sfc_desc* d, const void* in, void* out, char* info, int info_len) {
    sfc::PlanError err{0, ""};
    std::shared_ptr<sfc::Plan> p = sfc::Plan::create(*d, err);
    if (!p) {
        snprintf(info, (size_t)info_len, "%s", err.msg.c_str());
        return err.code ? err.code : -1;
    }
    std::string es;
    const int rc = p->exec(in, out, nullptr, es);
    snprintf(info, (size_t)info_len, "%s", rc == 0 ? p->describe().c_str() : es.c_str());
    return rc;
}

// the slab transpose fused into the store: output block q of the split axis goes to outs[q] (peer memory on the device)
extern "C" int emul_plan_run_scatter(const sfc_desc* d, const void* in, void* const* outs, int nouts, char* info, int info_len) {
    sfc::PlanError err{0, ""};
    std::shared_ptr<sfc::Plan> p = sfc::Plan::create(*d, err);
    if (!p) {
        snprintf(info, (size_t)info_len, "%s", err.msg.c_str());
        return err.code ? err.code : -1;
    }
    std::string es;
    const int rc = p->exec(in, nullptr, nullptr, es, outs, nouts);
    snprintf(info, (size_t)info_len, "%s", rc == 0 ? p->describe().c_str() : es.c_str());
    return rc;
}

// the same plan executed window by window (column blocks of every row of `row_lanes` lanes, last block first): the union of
// the windows must give what one full execution gives.  nouts == 0: plain output `out`; otherwise the scatter table.
extern "C" int emul_plan_run_windows(const sfc_desc* d, const void* in, void* out, void* const* outs, int nouts, long long row_lanes,
                                     int nwin, char* info, int info_len) {
    sfc::PlanError err{0, ""};
    std::shared_ptr<sfc::Plan> p = sfc::Plan::create(*d, err);
    if (!p) {
        snprintf(info, (size_t)info_len, "%s", err.msg.c_str());
        return err.code ? err.code : -1;
    }
    if (!p->window_ok(row_lanes, nwin)) {
        snprintf(info, (size_t)info_len, "window_ok() refused: %s", p->describe().c_str());
        return -2;
    }
    std::string es;
    for (int w = nwin - 1; w >= 0; --w) {
        sfc::ExecWindow win{row_lanes, w, nwin};
        const int rc = nouts ? p->exec(in, nullptr, nullptr, es, outs, nouts, &win) : p->exec(in, out, nullptr, es, nullptr, 0, &win);
        if (rc != 0) {
            snprintf(info, (size_t)info_len, "%s", es.c_str());
            return rc;
        }
    }
    snprintf(info, (size_t)info_len, "%s", p->describe().c_str());
    return 0;
}
