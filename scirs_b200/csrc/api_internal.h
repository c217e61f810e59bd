// api_internal.h — helpers shared by the C-ABI translation units (api.cu: the transforms of SURVEY 8a;
// api_ext.cu: the in-crate consumers of SURVEY 8f).  Nothing here is exported.
#pragma once
#include <cuda_runtime.h>
#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "plan.h"

namespace sfc_api {

int fail(int code, const std::string& msg);            // sets sfc_last_error(), returns code
int cuda_fail(cudaError_t e, const char* what);
std::shared_ptr<sfc::Plan> cached_plan(const sfc_desc& d, sfc::PlanError& err);  // through the global PlanCache
int ensure_buf(void** p, size_t* cap, size_t bytes);

inline bool dtype_is_complex(int dt) { return dt == SFC_C64 || dt == SFC_C128; }
inline bool dtype_is_f64(int dt) { return dt == SFC_F64 || dt == SFC_C128; }
inline size_t dtype_bytes(int dt) {
    switch (dt) {
        case SFC_F32: return 4;
        case SFC_F64: return 8;
        case SFC_C64: return 8;
        default: return 16;
    }
}
inline bool dtype_ok(int dt) { return dt >= SFC_F32 && dt <= SFC_C128; }

inline int64_t next_pow2_i64(int64_t n) {
    int64_t p = 1;
    while (p < n) p <<= 1;
    return p;
}

// NormMode / parse_norm_mode, fft/algorithms.rs:19-50
enum NormMode { NM_NONE, NM_BACKWARD, NM_ORTHO, NM_FORWARD };
inline NormMode parse_norm_mode(const char* norm, bool inverse) {
    if (!norm) return inverse ? NM_BACKWARD : NM_NONE;
    if (!strcmp(norm, "backward")) return NM_BACKWARD;
    if (!strcmp(norm, "ortho")) return NM_ORTHO;
    if (!strcmp(norm, "forward")) return NM_FORWARD;
    return NM_NONE;
}
// forward transforms: algorithms.rs:385-395 / 693-703 ; inverse: :528-534 / :876-884
inline double norm_scale(NormMode m, bool inverse, double total) {
    switch (m) {
        case NM_NONE: return 1.0;
        case NM_BACKWARD: return 1.0 / total;
        case NM_ORTHO: return 1.0 / std::sqrt(total);
        case NM_FORWARD: return inverse ? 1.0 : 1.0 / total;
    }
    return 1.0;
}

// per-thread device workspace for the host-pointer entry points
struct Workspace {
    void* buf[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t cap[6] = {0, 0, 0, 0, 0, 0};
    cudaStream_t stream = nullptr;
    int device = -1;
    ~Workspace() {
        for (int i = 0; i < 6; ++i)
            if (buf[i]) cudaFree(buf[i]);
        if (stream) cudaStreamDestroy(stream);
    }
    int get(int slot, size_t bytes, void** out) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev != device) {
            for (int i = 0; i < 6; ++i) {
                if (buf[i]) cudaFree(buf[i]);
                buf[i] = nullptr;
                cap[i] = 0;
            }
            if (stream) cudaStreamDestroy(stream);
            stream = nullptr;
            device = dev;
        }
        if (!stream) {
            cudaError_t e = cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
            if (e != cudaSuccess) return cuda_fail(e, "cudaStreamCreate");
        }
        int rc = ensure_buf(&buf[slot], &cap[slot], bytes);
        if (rc != SFC_OK) return rc;
        *out = buf[slot];
        return SFC_OK;
    }
};
extern thread_local Workspace g_ws;


int require_device();
int64_t vprod(const std::vector<int64_t>& v);
// upload x (in_shape, dtype), convert / pad / crop into complex f64 of tshape, transform over axes; result in *d_res
int run_c2c_host(const void* x, const std::vector<int64_t>& in_shape, int dtype, const std::vector<int64_t>& tshape,
                 const std::vector<int>& axes, bool inverse, double scale, void** d_res);
int download(void* h, const void* d, size_t bytes);    // D2H on the workspace stream + synchronize

// dist.cu: the free functions over several GPUs of this process (sfc_set_num_gpus); *handled says whether they ran
int multi_fftn_host(const void* x, const int64_t* shape3, const int* axes3, bool inverse, double scale, double* out, bool* handled);
int multi_batch_host(const sfc_desc& d, const void* in, void* out, bool* handled);

}  // namespace sfc_api
