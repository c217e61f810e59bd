// api.cu — the C ABI (include/scirs2_fft_cuda.h): plan cache, executor entry points and the
// drop-in free functions carrying the reference's wrapper semantics.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "api_internal.h"
#include "plan.h"

using namespace sfc;

// ------------------------------------------------------------------ errors

static thread_local std::string g_last_error;

namespace sfc {
void set_error(int code, const std::string& msg) {
    (void)code;
    g_last_error = msg;
}
}  // namespace sfc

namespace sfc_api {
int fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    cudaGetLastError();
    const int code = (e == cudaErrorMemoryAllocation) ? SFC_ERR_MEMORY : SFC_ERR_BACKEND;
    return fail(code, std::string(what) + ": " + cudaGetErrorString(e));
}
}  // namespace sfc_api
using namespace sfc_api;

// ------------------------------------------------------------- plan handle

struct sfc_plan {
    std::shared_ptr<Plan> p;
    // host-exec staging (device side), grown on demand
    void* d_in = nullptr;
    void* d_out = nullptr;
    size_t in_cap = 0, out_cap = 0;
    cudaStream_t stream = nullptr, s_h2d = nullptr, s_d2h = nullptr;
    std::vector<cudaEvent_t> ev;
    ~sfc_plan() {
        if (d_in) cudaFree(d_in);
        if (d_out) cudaFree(d_out);
        if (stream) cudaStreamDestroy(stream);
        if (s_h2d) cudaStreamDestroy(s_h2d);
        if (s_d2h) cudaStreamDestroy(s_d2h);
        for (cudaEvent_t e : ev) cudaEventDestroy(e);
    }
};

// --------------------------------------------------------------- plan cache
// plan_cache.rs:28-190: enabled flag, hit/miss counters (not touched while disabled),
// max_entries with LRU (last_used, usage_count) eviction, max_age TTL.

namespace {

struct CacheKey {
    std::string bytes;
    bool operator<(const CacheKey& o) const { return bytes < o.bytes; }
};

struct CacheEntry {
    std::shared_ptr<Plan> plan;
    std::chrono::steady_clock::time_point last_used;
    uint64_t usage_count;
};

struct PlanCacheImpl {
    std::mutex mu;
    std::map<CacheKey, CacheEntry> map;
    uint64_t max_entries = 128;
    double max_age_s = 3600.0;
    bool enabled = true;
    uint64_t hits = 0, misses = 0;
};

PlanCacheImpl& cache() {
    static PlanCacheImpl c;
    return c;
}

CacheKey make_key(const sfc_desc& d) {
    sfc_desc k;
    memset(&k, 0, sizeof k);
    k.ndim = d.ndim;
    for (int i = 0; i < d.ndim && i < SFC_MAX_DIMS; ++i) k.shape[i] = d.shape[i];
    k.naxes = d.naxes;
    for (int i = 0; i < d.naxes && i < SFC_MAX_DIMS; ++i) k.axes[i] = d.axes[i];
    k.kind = d.kind;
    k.prec = d.prec;
    k.direction = d.kind == SFC_C2C ? d.direction : 0;
    k.flags = d.flags;
    k.scale = d.scale;
    k.scatter_parts = d.scatter_parts;
    k.scatter_pitch = d.scatter_parts > 1 ? d.scatter_pitch : 0;
    if (d.flags & SFC_DESC_AXIS_LEN) {
        k.axis_in_len = d.axis_in_len;
        k.axis_out_len = d.axis_out_len;
    }
    if (d.flags & SFC_DESC_AUX_MUL) {
        k.aux_in = d.aux_in;
        k.aux_out = d.aux_out;
    }
    if (d.flags & (SFC_DESC_DCT2 | SFC_DESC_DCT3)) k.scale_dc = d.scale_dc;
    if (d.flags & SFC_DESC_CUSTOM_IN_SHAPE)
        for (int i = 0; i < d.ndim && i < SFC_MAX_DIMS; ++i) k.in_shape[i] = d.in_shape[i];
    int dev = 0;
    cudaGetDevice(&dev);
    CacheKey ck;
    ck.bytes.assign(reinterpret_cast<const char*>(&k), sizeof k);
    ck.bytes.append(reinterpret_cast<const char*>(&dev), sizeof dev);
    return ck;
}

void evict_old_entries(PlanCacheImpl& c) {
    const auto now = std::chrono::steady_clock::now();
    for (auto it = c.map.begin(); it != c.map.end();) {
        const double age = std::chrono::duration<double>(now - it->second.last_used).count();
        if (age > c.max_age_s)
            it = c.map.erase(it);
        else
            ++it;
    }
    while (!c.map.empty() && c.map.size() >= c.max_entries) {
        auto victim = c.map.begin();
        for (auto it = c.map.begin(); it != c.map.end(); ++it) {
            if (std::make_pair(it->second.last_used, it->second.usage_count) <
                std::make_pair(victim->second.last_used, victim->second.usage_count))
                victim = it;
        }
        c.map.erase(victim);
    }
}

// device-memory pressure (plan.h): give back the scratch of every cached plan that is idle
size_t release_cached_scratch(const Plan* except) {
    PlanCacheImpl& c = cache();
    std::vector<std::shared_ptr<Plan>> plans;
    {
        std::lock_guard<std::mutex> lk(c.mu);
        for (auto& kv : c.map) plans.push_back(kv.second.plan);
    }
    size_t freed = 0;
    for (auto& p : plans)
        if (p.get() != except) freed += p->release_scratch();
    return freed;
}

std::shared_ptr<Plan> get_or_create_plan(const sfc_desc& d, PlanError& err) {
    PlanCacheImpl& c = cache();
    static const bool hooked = (sfc::set_scratch_pressure_hook(&release_cached_scratch), true);
    (void)hooked;
    {
        std::lock_guard<std::mutex> lk(c.mu);
        if (!c.enabled) {
            // plan_cache.rs:108-114: bypass, counters untouched
        } else {
            const CacheKey key = make_key(d);
            auto it = c.map.find(key);
            if (it != c.map.end()) {
                const double age =
                    std::chrono::duration<double>(std::chrono::steady_clock::now() - it->second.last_used).count();
                if (age <= c.max_age_s) {
                    it->second.last_used = std::chrono::steady_clock::now();
                    it->second.usage_count += 1;
                    c.hits += 1;
                    return it->second.plan;
                }
                c.map.erase(it);
            }
            c.misses += 1;
        }
    }
    std::shared_ptr<Plan> p = Plan::create(d, err);
    if (!p) return nullptr;
    std::lock_guard<std::mutex> lk(c.mu);
    if (c.enabled && c.max_entries > 0) {
        if (c.map.size() >= c.max_entries) evict_old_entries(c);
        c.map[make_key(d)] = CacheEntry{p, std::chrono::steady_clock::now(), 1};
    }
    return p;
}

}  // namespace

namespace sfc_api {
std::shared_ptr<Plan> cached_plan(const sfc_desc& d, PlanError& err) { return get_or_create_plan(d, err); }
}  // namespace sfc_api

// ------------------------------------------------------------------ runtime

extern "C" __attribute__((visibility("default"))) int sfc_abi_version(void) { return SFC_ABI_VERSION; }

extern "C" __attribute__((visibility("default"))) const char* sfc_last_error(void) { return g_last_error.c_str(); }

extern "C" __attribute__((visibility("default"))) int sfc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" __attribute__((visibility("default"))) int sfc_is_available(void) { return sfc_device_count() > 0 ? 1 : 0; }

extern "C" __attribute__((visibility("default"))) int sfc_init(int device) {
    const int n = sfc_device_count();
    if (n == 0) return fail(SFC_ERR_BACKEND, "no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= n) return fail(SFC_ERR_VALUE, "device index out of range");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    return SFC_OK;
}

// -------------------------------------------------------------------- plans

extern "C" __attribute__((visibility("default"))) int sfc_plan_create(sfc_plan** out, const sfc_desc* desc) {
    if (!out || !desc) return fail(SFC_ERR_VALUE, "null argument");
    *out = nullptr;
    PlanError err{0, ""};
    std::shared_ptr<Plan> p = get_or_create_plan(*desc, err);
    if (!p) return fail(err.code ? err.code : SFC_ERR_PLAN, err.msg);
    sfc_plan* h = new sfc_plan();
    h->p = p;
    *out = h;
    return SFC_OK;
}

extern "C" __attribute__((visibility("default"))) int sfc_plan_destroy(sfc_plan* plan) {
    delete plan;
    return SFC_OK;
}

extern "C" __attribute__((visibility("default"))) int sfc_plan_get_info(const sfc_plan* plan, sfc_plan_info* info) {
    if (!plan || !info) return fail(SFC_ERR_VALUE, "null argument");
    *info = plan->p->info;
    return SFC_OK;
}

extern "C" __attribute__((visibility("default"))) int sfc_plan_describe(const sfc_plan* plan, char* buf, size_t cap) {
    if (!plan || !buf || cap == 0) return fail(SFC_ERR_VALUE, "null argument");
    const std::string s = plan->p->describe();
    const size_t n = std::min(cap - 1, s.size());
    memcpy(buf, s.data(), n);
    buf[n] = 0;
    return (int)n;
}

extern "C" __attribute__((visibility("default"))) int sfc_exec_device(sfc_plan* plan, const void* d_in, void* d_out, void* stream) {
    if (!plan || !d_in || !d_out) return fail(SFC_ERR_VALUE, "null argument");
    std::string es;
    const int rc = plan->p->exec(d_in, d_out, (cudaStream_t)stream, es);
    if (rc != 0) return fail(rc, es);
    return SFC_OK;
}

extern "C" __attribute__((visibility("default"))) int sfc_exec_device_scatter(sfc_plan* plan, const void* d_in,
                                                                             void* const* d_outs, int32_t nouts,
                                                                             void* stream) {
    if (!plan || !d_in || !d_outs || nouts <= 0) return fail(SFC_ERR_VALUE, "null argument");
    std::string es;
    const int rc = plan->p->exec(d_in, d_outs[0], (cudaStream_t)stream, es, d_outs, nouts);
    if (rc != 0) return fail(rc, es);
    return SFC_OK;
}

// ------------------------------------------------------------ peer memory

extern "C" __attribute__((visibility("default"))) int sfc_dev_malloc(void** d_ptr, size_t bytes) {
    if (!d_ptr) return fail(SFC_ERR_VALUE, "null argument");
    cudaError_t e = sfc::alloc_with_relief(d_ptr, bytes ? bytes : 256, nullptr);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    return SFC_OK;
}

extern "C" __attribute__((visibility("default"))) int sfc_dev_free(void* d_ptr) {
    cudaError_t e = cudaFree(d_ptr);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFree");
    return SFC_OK;
}

extern "C" __attribute__((visibility("default"))) int sfc_ipc_get_handle(const void* d_ptr, void* handle_out) {
    if (!d_ptr || !handle_out) return fail(SFC_ERR_VALUE, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == SFC_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void*>(d_ptr));
    if (e != cudaSuccess) return cuda_fail(e, "cudaIpcGetMemHandle");
    memcpy(handle_out, &h, sizeof h);
    return SFC_OK;
}

extern "C" __attribute__((visibility("default"))) int sfc_ipc_open_handle(const void* handle, void** d_ptr_out) {
    if (!handle || !d_ptr_out) return fail(SFC_ERR_VALUE, "null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    cudaError_t e = cudaIpcOpenMemHandle(d_ptr_out, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(SFC_ERR_COMMUNICATION, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
    }
    return SFC_OK;
}

extern "C" __attribute__((visibility("default"))) int sfc_ipc_close_handle(void* d_ptr) {
    cudaError_t e = cudaIpcCloseMemHandle(d_ptr);
    if (e != cudaSuccess) return cuda_fail(e, "cudaIpcCloseMemHandle");
    return SFC_OK;
}

extern "C" __attribute__((visibility("default"))) int sfc_stream_synchronize(void* stream) {
    cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaStreamSynchronize");
    return SFC_OK;
}

int sfc_api::ensure_buf(void** p, size_t* cap, size_t bytes) {
    if (*cap >= bytes && *p) return SFC_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    cudaError_t e = sfc::alloc_with_relief(p, std::max<size_t>(bytes, 256), nullptr);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    *cap = std::max<size_t>(bytes, 256);
    return SFC_OK;
}

// Large batched plans are executed as a 3-stream pipeline over chunks of the outermost
// (untransformed) dimension: H2D of chunk i+1, kernels of chunk i and D2H of chunk i-1 overlap, so the
// call costs ~max(H2D, D2H) instead of their sum (PCIe is full duplex).
static int exec_host_pipelined(sfc_plan* plan, const void* h_in, void* h_out, int nchunks) {
    Plan& p = *plan->p;
    const int64_t n0 = p.desc.shape[0];
    const size_t in_row = (size_t)p.info.in_bytes / (size_t)n0, out_row = (size_t)p.info.out_bytes / (size_t)n0;
    if (!plan->s_h2d) {
        cudaError_t e = cudaStreamCreateWithFlags(&plan->s_h2d, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&plan->s_d2h, cudaStreamNonBlocking);
        if (e != cudaSuccess) return cuda_fail(e, "cudaStreamCreate");
    }
    if ((int)plan->ev.size() < 2 * nchunks) {
        const size_t old = plan->ev.size();
        plan->ev.resize(2 * nchunks);
        for (size_t i = old; i < plan->ev.size(); ++i) {
            cudaError_t e = cudaEventCreateWithFlags(&plan->ev[i], cudaEventDisableTiming);
            if (e != cudaSuccess) return cuda_fail(e, "cudaEventCreate");
        }
    }
    const int64_t per = (n0 + nchunks - 1) / nchunks;
    int ci = 0;
    for (int64_t r0 = 0; r0 < n0; r0 += per, ++ci) {
        const int64_t rows = std::min(per, n0 - r0);
        sfc_desc d = p.desc;
        d.shape[0] = rows;
        if (d.flags & SFC_DESC_CUSTOM_IN_SHAPE) d.in_shape[0] = rows;
        PlanError perr{0, ""};
        std::shared_ptr<Plan> sub = (rows == n0) ? plan->p : get_or_create_plan(d, perr);
        if (!sub) return fail(perr.code ? perr.code : SFC_ERR_PLAN, perr.msg);
        char* din = (char*)plan->d_in + (size_t)r0 * in_row;
        char* dout = (char*)plan->d_out + (size_t)r0 * out_row;
        cudaError_t e = cudaMemcpyAsync(din, (const char*)h_in + (size_t)r0 * in_row, (size_t)rows * in_row,
                                        cudaMemcpyHostToDevice, plan->s_h2d);
        if (e != cudaSuccess) return cuda_fail(e, "H2D copy");
        cudaEventRecord(plan->ev[2 * ci], plan->s_h2d);
        cudaStreamWaitEvent(plan->stream, plan->ev[2 * ci], 0);
        std::string es;
        int rc = sub->exec(din, dout, plan->stream, es);
        if (rc != 0) return fail(rc, es);
        cudaEventRecord(plan->ev[2 * ci + 1], plan->stream);
        cudaStreamWaitEvent(plan->s_d2h, plan->ev[2 * ci + 1], 0);
        e = cudaMemcpyAsync((char*)h_out + (size_t)r0 * out_row, dout, (size_t)rows * out_row, cudaMemcpyDeviceToHost,
                            plan->s_d2h);
        if (e != cudaSuccess) return cuda_fail(e, "D2H copy");
    }
    cudaError_t e = cudaStreamSynchronize(plan->s_d2h);
    if (e == cudaSuccess) e = cudaStreamSynchronize(plan->stream);
    if (e != cudaSuccess) return cuda_fail(e, "transform execution");
    return SFC_OK;
}

extern "C" __attribute__((visibility("default"))) int sfc_exec_host(sfc_plan* plan, const void* h_in, void* h_out) {
    if (!plan || !h_in || !h_out) return fail(SFC_ERR_VALUE, "null argument");
    Plan& p = *plan->p;
    int rc;
    if ((rc = ensure_buf(&plan->d_in, &plan->in_cap, (size_t)p.info.in_bytes)) != SFC_OK) return rc;
    if ((rc = ensure_buf(&plan->d_out, &plan->out_cap, (size_t)p.info.out_bytes)) != SFC_OK) return rc;
    if (!plan->stream) {
        cudaError_t e = cudaStreamCreateWithFlags(&plan->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) return cuda_fail(e, "cudaStreamCreate");
    }
    // pipeline when dimension 0 is a pure batch dimension and the transfer is big enough to matter
    {
        bool batch0 = p.desc.ndim >= 2 && p.desc.shape[0] >= 2 && p.desc.scatter_parts <= 1;
        for (int i = 0; i < p.desc.naxes; ++i) batch0 = batch0 && p.desc.axes[i] != 0;
        const int64_t bytes = p.info.in_bytes + p.info.out_bytes;
        static const int want = [] {
            const char* e = getenv("SFC_HOST_CHUNKS");
            return e ? atoi(e) : 16;  // fill + drain cost 1/chunks of a one-way transfer
        }();
        if (batch0 && want > 1 && bytes >= ((int64_t)64 << 20)) {
            const int nchunks = (int)std::min<int64_t>(want, p.desc.shape[0]);
            return exec_host_pipelined(plan, h_in, h_out, nchunks);
        }
    }
    cudaStream_t s = plan->stream;
    cudaError_t e = cudaMemcpyAsync(plan->d_in, h_in, (size_t)p.info.in_bytes, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return cuda_fail(e, "H2D copy");
    std::string es;
    rc = p.exec(plan->d_in, plan->d_out, s, es);
    if (rc != 0) return fail(rc, es);
    e = cudaMemcpyAsync(h_out, plan->d_out, (size_t)p.info.out_bytes, cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) return cuda_fail(e, "D2H copy");
    e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return cuda_fail(e, "transform execution");
    return SFC_OK;
}

// --------------------------------------------------------------- plan cache

extern "C" __attribute__((visibility("default"))) int sfc_cache_get_stats(sfc_cache_stats* out) {
    if (!out) return fail(SFC_ERR_VALUE, "null argument");
    PlanCacheImpl& c = cache();
    std::lock_guard<std::mutex> lk(c.mu);
    out->hit_count = c.hits;
    out->miss_count = c.misses;
    const uint64_t tot = c.hits + c.misses;
    out->hit_rate = tot ? (double)c.hits / (double)tot : 0.0;
    out->size = c.map.size();
    out->max_size = c.max_entries;
    return SFC_OK;
}

extern "C" __attribute__((visibility("default"))) int sfc_cache_set_enabled(int enabled) {
    PlanCacheImpl& c = cache();
    std::lock_guard<std::mutex> lk(c.mu);
    c.enabled = enabled != 0;
    return SFC_OK;
}

extern "C" __attribute__((visibility("default"))) int sfc_cache_is_enabled(void) {
    PlanCacheImpl& c = cache();
    std::lock_guard<std::mutex> lk(c.mu);
    return c.enabled ? 1 : 0;
}

extern "C" __attribute__((visibility("default"))) int sfc_cache_clear(void) {
    PlanCacheImpl& c = cache();
    std::lock_guard<std::mutex> lk(c.mu);
    c.map.clear();
    return SFC_OK;
}

extern "C" __attribute__((visibility("default"))) int sfc_cache_configure(uint64_t max_entries, double max_age_seconds) {
    PlanCacheImpl& c = cache();
    std::lock_guard<std::mutex> lk(c.mu);
    // PlanCache::with_config builds a fresh cache: entries and counters reset
    c.map.clear();
    c.hits = c.misses = 0;
    c.max_entries = max_entries;
    c.max_age_s = max_age_seconds;
    return SFC_OK;
}

// ---------------------------------------------------------- planner options

extern "C" __attribute__((visibility("default"))) int sfc_planner_set_option(const char* name, const char* value) {
    if (!name || strncmp(name, "SFC_", 4) != 0 || strlen(name) > 64) return fail(SFC_ERR_VALUE, "option names start with SFC_");
    sfc::planner_set_option(name, value);
    // cached plans were built under the old options
    PlanCacheImpl& c = cache();
    std::lock_guard<std::mutex> lk(c.mu);
    c.map.clear();
    return SFC_OK;
}

extern "C" __attribute__((visibility("default"))) int sfc_planner_get_option(const char* name, char* buf, size_t cap) {
    if (!name || !buf || cap == 0) return fail(SFC_ERR_VALUE, "null argument");
    const std::string v = sfc::planner_get_option(name);
    const size_t n = std::min(cap - 1, v.size());
    memcpy(buf, v.data(), n);
    buf[n] = 0;
    return (int)n;
}

// ====================================================== drop-in free functions

namespace sfc_api {

thread_local Workspace g_ws;

int require_device() {
    if (sfc_device_count() == 0)
        return fail(SFC_ERR_BACKEND, "no CUDA device available (this library has no CPU fallback)");
    return SFC_OK;
}

int64_t vprod(const std::vector<int64_t>& v) {
    int64_t p = 1;
    for (int64_t x : v) p *= x;
    return p;
}

// Upload `x` (in_shape, dtype), convert / pad / crop into a complex f64 array of `tshape`
// and run a complex transform over `axes` on it; result (complex f64, tshape) left in *d_res.
// `x_host` may be replaced by `x_host_override` already being complex f64 of tshape.
int run_c2c_host(const void* x, const std::vector<int64_t>& in_shape, int dtype, const std::vector<int64_t>& tshape,
                 const std::vector<int>& axes, bool inverse, double scale, void** d_res) {
    int rc;
    const int64_t in_total = vprod(in_shape), t_total = vprod(tshape);
    void *d_raw = nullptr, *d_work = nullptr;
    if ((rc = g_ws.get(0, (size_t)in_total * dtype_bytes(dtype), &d_raw)) != SFC_OK) return rc;
    if ((rc = g_ws.get(1, (size_t)t_total * 16, &d_work)) != SFC_OK) return rc;
    cudaStream_t s = g_ws.stream;
    cudaError_t e = cudaMemcpyAsync(d_raw, x, (size_t)in_total * dtype_bytes(dtype), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return cuda_fail(e, "H2D copy");

    sfc_desc d;
    memset(&d, 0, sizeof d);
    d.ndim = (int)tshape.size();
    for (int i = 0; i < d.ndim; ++i) d.shape[i] = tshape[i];
    d.naxes = (int)axes.size();
    for (int i = 0; i < d.naxes; ++i) d.axes[i] = axes[i];
    d.kind = SFC_C2C;
    d.prec = SFC_PREC_F64;
    d.direction = inverse ? SFC_INVERSE : SFC_FORWARD;
    d.scale = scale;

    const bool same = (in_shape == tshape);
    const void* src = d_raw;
    if (same && dtype == SFC_C128) {
        // transform straight out of the upload buffer
    } else if (same && dtype == SFC_F64 && !axes.empty()) {
        d.flags |= SFC_DESC_REAL_INPUT;
    } else {
        // convert_to_complex + pad/crop (algorithms.rs:617-664) as one device pass
        CopyParams c;
        memset(&c, 0, sizeof c);
        c.ndim = (int)tshape.size();
        for (int i = 0; i < c.ndim; ++i) {
            c.dst_shape[i] = tshape[i];
            c.src_shape[i] = in_shape[i];
        }
        c.src_complex = dtype_is_complex(dtype);
        c.dst_complex = 1;
        c.src_f64 = dtype_is_f64(dtype);
        c.dst_f64 = 1;
        c.scale = 1.0;
        c.total = t_total;
        c.src = d_raw;
        c.dst = d_work;
        e = launch_nd_copy(c, s);
        if (e != cudaSuccess) return cuda_fail(e, "convert/pad kernel");
        src = d_work;
    }
    PlanError perr{0, ""};
    std::shared_ptr<Plan> p = cached_plan(d, perr);
    if (!p) return fail(perr.code ? perr.code : SFC_ERR_PLAN, perr.msg);
    std::string es;
    rc = p->exec(src, d_work, s, es);
    if (rc != 0) return fail(rc, es);
    *d_res = d_work;
    return SFC_OK;
}

int download(void* h, const void* d, size_t bytes) {
    cudaStream_t s = g_ws.stream;
    cudaError_t e = cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) return cuda_fail(e, "D2H copy");
    e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return cuda_fail(e, "transform execution");
    return SFC_OK;
}

}  // namespace sfc_api

namespace {

int fft1d_common(const void* x, int64_t len, int dtype, int64_t n, bool inverse, double* out, int64_t out_cap,
                 int64_t* out_len) {
    int rc;
    if ((rc = require_device()) != SFC_OK) return rc;
    if (!dtype_ok(dtype)) return fail(SFC_ERR_VALUE, "unknown dtype");
    if (len <= 0 || !x) return fail(SFC_ERR_VALUE, "Input cannot be empty");  // algorithms.rs:136-138
    if (n == 0) return fail(SFC_ERR_VALUE, "FFT size must be positive");
    const int64_t fft_size = n > 0 ? n : next_pow2_i64(len);  // algorithms.rs:142 (pads to next pow2!)
    int64_t produced = fft_size;
    if (inverse && n <= 0 && fft_size > len) produced = len;  // algorithms.rs:258-260
    if (out_len) *out_len = produced;
    if (!out || out_cap < produced) return fail(SFC_ERR_VALUE, "output buffer too small");
    void* d_res = nullptr;
    const double scale = inverse ? 1.0 / (double)fft_size : 1.0;  // algorithms.rs:255
    rc = run_c2c_host(x, {len}, dtype, {fft_size}, {0}, inverse, scale, &d_res);
    if (rc != SFC_OK) return rc;
    return download(out, d_res, (size_t)produced * 16);
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int sfc_fft(const void* x, int64_t len, int dtype, int64_t n, double* out, int64_t out_cap,
                       int64_t* out_len) {
    return fft1d_common(x, len, dtype, n, false, out, out_cap, out_len);
}

extern "C" __attribute__((visibility("default"))) int sfc_ifft(const void* x, int64_t len, int dtype, int64_t n, double* out, int64_t out_cap,
                        int64_t* out_len) {
    return fft1d_common(x, len, dtype, n, true, out, out_cap, out_len);
}

extern "C" __attribute__((visibility("default"))) int sfc_rfft(const void* x, int64_t len, int dtype, int64_t n, double* out, int64_t out_cap,
                        int64_t* out_len) {
    int rc;
    if ((rc = require_device()) != SFC_OK) return rc;
    if (!dtype_ok(dtype)) return fail(SFC_ERR_VALUE, "unknown dtype");
    if (len <= 0 || !x) return fail(SFC_ERR_VALUE, "Input cannot be empty");
    if (n == 0) return fail(SFC_ERR_VALUE, "FFT size must be positive");
    const int64_t n_val = n > 0 ? n : len;  // rfft.rs:45 (no pow2 padding here)
    const int64_t n_out = n_val / 2 + 1;    // rfft.rs:51
    if (out_len) *out_len = n_out;
    if (!out || out_cap < n_out) return fail(SFC_ERR_VALUE, "output buffer too small");
    if (dtype_is_complex(dtype)) {
        // complex input: the reference runs the full fft and keeps the first n/2+1 bins
        void* d_res = nullptr;
        rc = run_c2c_host(x, {len}, dtype, {n_val}, {0}, false, 1.0, &d_res);
        if (rc != SFC_OK) return rc;
        return download(out, d_res, (size_t)n_out * 16);
    }
    // real input: upload, widen / pad to f64[n_val], real->complex plan
    void *d_raw = nullptr, *d_real = nullptr, *d_outb = nullptr;
    if ((rc = g_ws.get(0, (size_t)len * dtype_bytes(dtype), &d_raw)) != SFC_OK) return rc;
    if ((rc = g_ws.get(1, (size_t)n_val * 8, &d_real)) != SFC_OK) return rc;
    if ((rc = g_ws.get(2, (size_t)n_out * 16, &d_outb)) != SFC_OK) return rc;
    cudaStream_t s = g_ws.stream;
    cudaError_t e = cudaMemcpyAsync(d_raw, x, (size_t)len * dtype_bytes(dtype), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return cuda_fail(e, "H2D copy");
    const void* src = d_raw;
    if (!(dtype == SFC_F64 && len == n_val)) {
        CopyParams c;
        memset(&c, 0, sizeof c);
        c.ndim = 1;
        c.dst_shape[0] = n_val;
        c.src_shape[0] = len;
        c.src_complex = 0;
        c.dst_complex = 0;
        c.src_f64 = dtype_is_f64(dtype);
        c.dst_f64 = 1;
        c.scale = 1.0;
        c.total = n_val;
        c.src = d_raw;
        c.dst = d_real;
        e = launch_nd_copy(c, s);
        if (e != cudaSuccess) return cuda_fail(e, "convert/pad kernel");
        src = d_real;
    }
    sfc_desc d;
    memset(&d, 0, sizeof d);
    d.ndim = 1;
    d.shape[0] = n_val;
    d.naxes = 1;
    d.axes[0] = 0;
    d.kind = SFC_R2C;
    d.prec = SFC_PREC_F64;
    d.scale = 1.0;
    PlanError perr{0, ""};
    std::shared_ptr<Plan> p = cached_plan(d, perr);
    if (!p) return fail(perr.code ? perr.code : SFC_ERR_PLAN, perr.msg);
    std::string es;
    rc = p->exec(src, d_outb, s, es);
    if (rc != 0) return fail(rc, es);
    return download(out, d_outb, (size_t)n_out * 16);
}

extern "C" __attribute__((visibility("default"))) int sfc_irfft(const void* x, int64_t len, int dtype, int64_t n, double* out, int64_t out_cap,
                         int64_t* out_len) {
    int rc;
    if ((rc = require_device()) != SFC_OK) return rc;
    if (!dtype_ok(dtype)) return fail(SFC_ERR_VALUE, "unknown dtype");
    if (len <= 0 || !x) return fail(SFC_ERR_VALUE, "Input cannot be empty");
    if (n == 0) return fail(SFC_ERR_VALUE, "FFT size must be positive");
    const int64_t n_output = n > 0 ? n : 2 * (len - 1);  // rfft.rs:138-141
    if (n_output <= 0) return fail(SFC_ERR_VALUE, "Input cannot be empty");  // ifft of an empty spectrum
    if (out_len) *out_len = n_output;
    if (!out || out_cap < n_output) return fail(SFC_ERR_VALUE, "output buffer too small");

    // widen the input to Complex64 on the host (tiny: len elements)
    std::vector<double> xin(2 * (size_t)len);
    for (int64_t i = 0; i < len; ++i) {
        double re, im = 0.0;
        switch (dtype) {
            case SFC_F32: re = ((const float*)x)[i]; break;
            case SFC_F64: re = ((const double*)x)[i]; break;
            case SFC_C64: re = ((const float*)x)[2 * i]; im = ((const float*)x)[2 * i + 1]; break;
            default: re = ((const double*)x)[2 * i]; im = ((const double*)x)[2 * i + 1]; break;
        }
        xin[2 * i] = re;
        xin[2 * i + 1] = im;
    }
    cudaStream_t s;
    const bool fast = (n_output % 2 == 0) && (len == n_output / 2 + 1);
    if (fast) {
        // proper half spectrum: complex->real plan (fused pre-twiddle when n is a power of two)
        void *d_in = nullptr, *d_o = nullptr;
        if ((rc = g_ws.get(0, (size_t)len * 16, &d_in)) != SFC_OK) return rc;
        if ((rc = g_ws.get(1, (size_t)n_output * 8, &d_o)) != SFC_OK) return rc;
        s = g_ws.stream;
        cudaError_t e = cudaMemcpyAsync(d_in, xin.data(), (size_t)len * 16, cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) return cuda_fail(e, "H2D copy");
        sfc_desc d;
        memset(&d, 0, sizeof d);
        d.ndim = 1;
        d.shape[0] = n_output;
        d.naxes = 1;
        d.axes[0] = 0;
        d.kind = SFC_C2R;
        d.prec = SFC_PREC_F64;
        d.scale = 1.0 / (double)n_output;
        PlanError perr{0, ""};
        std::shared_ptr<Plan> p = cached_plan(d, perr);
        if (!p) return fail(perr.code ? perr.code : SFC_ERR_PLAN, perr.msg);
        std::string es;
        rc = p->exec(d_in, d_o, s, es);
        if (rc != 0) return fail(rc, es);
        return download(out, d_o, (size_t)n_output * 8);
    }
    // literal reference path (rfft.rs:143-175): extend, truncating complex ifft, real part
    std::vector<double> full;
    full.reserve(2 * (size_t)std::max(n_output, len));
    full.assign(xin.begin(), xin.end());
    if (n_output > len) {
        const int64_t start_idx = (n_output % 2 == 0) ? len - 1 : len;
        for (int64_t i = start_idx - 1; i >= 1; --i) {
            if ((int64_t)(full.size() / 2) >= n_output) break;
            full.push_back(xin[2 * i]);
            full.push_back(-xin[2 * i + 1]);
        }
        full.resize(2 * (size_t)n_output, 0.0);
    }
    const int64_t flen = (int64_t)(full.size() / 2);
    void* d_res = nullptr;
    rc = run_c2c_host(full.data(), {flen}, SFC_C128, {n_output}, {0}, true, 1.0 / (double)n_output, &d_res);
    if (rc != SFC_OK) return rc;
    std::vector<double> tmp(2 * (size_t)n_output);
    rc = download(tmp.data(), d_res, (size_t)n_output * 16);
    if (rc != SFC_OK) return rc;
    for (int64_t i = 0; i < n_output; ++i) out[i] = tmp[2 * i];
    return SFC_OK;
}

namespace {

int fft2_common(const void* x, int64_t rows, int64_t cols, int dtype, const int64_t* shape2, const int32_t* axes2,
                const char* norm, bool inverse, double* out, int64_t out_cap, int64_t* out_shape2) {
    int rc;
    if ((rc = require_device()) != SFC_OK) return rc;
    if (!dtype_ok(dtype)) return fail(SFC_ERR_VALUE, "unknown dtype");
    if (!x || rows <= 0 || cols <= 0) return fail(SFC_ERR_VALUE, "Input cannot be empty");
    const int64_t o0 = shape2 ? shape2[0] : rows, o1 = shape2 ? shape2[1] : cols;
    if (o0 <= 0 || o1 <= 0) return fail(SFC_ERR_VALUE, "FFT size must be positive");
    const int a0 = axes2 ? axes2[0] : 0, a1 = axes2 ? axes2[1] : 1;
    // validated, then ignored (algorithms.rs:309-314): rows are always transformed first
    if (a0 < 0 || a0 > 1 || a1 < 0 || a1 > 1 || a0 == a1)
        return fail(SFC_ERR_VALUE, inverse ? "Invalid axes for 2D IFFT" : "Invalid axes for 2D FFT");
    if (out_shape2) {
        out_shape2[0] = o0;
        out_shape2[1] = o1;
    }
    if (!out || out_cap < o0 * o1) return fail(SFC_ERR_VALUE, "output buffer too small");
    const double scale = norm_scale(parse_norm_mode(norm, inverse), inverse, (double)o0 * (double)o1);
    void* d_res = nullptr;
    rc = run_c2c_host(x, {rows, cols}, dtype, {o0, o1}, {1, 0}, inverse, scale, &d_res);
    if (rc != SFC_OK) return rc;
    return download(out, d_res, (size_t)(o0 * o1) * 16);
}

int fftn_common(const void* x, int32_t ndim, const int64_t* in_shape, int dtype, const int64_t* shape,
                const int64_t* axes, int32_t naxes, const char* norm, bool inverse, double* out, int64_t out_cap,
                int64_t* out_shape, std::vector<int64_t>* tshape_out, std::vector<int>* axes_out, void** d_keep) {
    int rc;
    if ((rc = require_device()) != SFC_OK) return rc;
    if (!dtype_ok(dtype)) return fail(SFC_ERR_VALUE, "unknown dtype");
    if (ndim < 1 || ndim > SFC_MAX_DIMS) return fail(SFC_ERR_VALUE, "ndim must be in 1..8");
    if (!x || !in_shape) return fail(SFC_ERR_VALUE, "Input cannot be empty");
    std::vector<int64_t> ish(in_shape, in_shape + ndim);
    for (int64_t v : ish)
        if (v <= 0) return fail(SFC_ERR_VALUE, "Input cannot be empty");
    // `shape` must have ndim entries (algorithms.rs:594-598); callers pass nullptr for None
    std::vector<int64_t> tsh = shape ? std::vector<int64_t>(shape, shape + ndim) : ish;
    for (int64_t v : tsh)
        if (v <= 0) return fail(SFC_ERR_VALUE, "FFT size must be positive");
    std::vector<int> ax;
    if (axes) {
        for (int i = 0; i < naxes; ++i) {
            if (axes[i] < 0 || axes[i] >= ndim) {  // algorithms.rs:604-611
                char b[96];
                snprintf(b, sizeof b, "Axis %lld out of bounds for array of dimension %d", (long long)axes[i], ndim);
                return fail(SFC_ERR_VALUE, b);
            }
            ax.push_back((int)axes[i]);
        }
    } else {
        for (int i = 0; i < ndim; ++i) ax.push_back(i);
    }
    if ((int)ax.size() > SFC_MAX_DIMS) return fail(SFC_ERR_VALUE, "too many axes");
    const NormMode nm = parse_norm_mode(norm, inverse);
    double total;
    if (!inverse) {
        total = (double)vprod(tsh);  // ALL dims, not just the transformed ones (algorithms.rs:694)
    } else {
        total = 1.0;  // only the listed axes (algorithms.rs:876), duplicates counted twice
        for (int a : ax) total *= (double)tsh[a];
    }
    const double scale = norm_scale(nm, inverse, total);
    if (out_shape)
        for (int i = 0; i < ndim; ++i) out_shape[i] = tsh[i];
    if (tshape_out) *tshape_out = tsh;
    if (axes_out) *axes_out = ax;
    void* d_res = nullptr;
    if (d_keep) {
        rc = run_c2c_host(x, ish, dtype, tsh, ax, inverse, scale, &d_res);
        if (rc != SFC_OK) return rc;
        *d_keep = d_res;
        return SFC_OK;
    }
    if (!out || out_cap < vprod(tsh)) return fail(SFC_ERR_VALUE, "output buffer too small");
    if (ndim == 3 && dtype == SFC_C128 && ish == tsh && ax.size() == 3 && ax[0] != ax[1] && ax[0] != ax[2] && ax[1] != ax[2]) {
        // sfc_set_num_gpus(P > 1): slab decomposition over the GPUs of this process (dist.cu)
        bool handled = false;
        rc = multi_fftn_host(x, tsh.data(), ax.data(), inverse, scale, out, &handled);
        if (rc != SFC_OK || handled) return rc;
    }
    rc = run_c2c_host(x, ish, dtype, tsh, ax, inverse, scale, &d_res);
    if (rc != SFC_OK) return rc;
    return download(out, d_res, (size_t)vprod(tsh) * 16);
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int sfc_fft2(const void* x, int64_t rows, int64_t cols, int dtype, const int64_t* shape2,
                        const int32_t* axes2, const char* norm, double* out, int64_t out_cap, int64_t* out_shape2) {
    return fft2_common(x, rows, cols, dtype, shape2, axes2, norm, false, out, out_cap, out_shape2);
}

extern "C" __attribute__((visibility("default"))) int sfc_ifft2(const void* x, int64_t rows, int64_t cols, int dtype, const int64_t* shape2,
                         const int32_t* axes2, const char* norm, double* out, int64_t out_cap, int64_t* out_shape2) {
    return fft2_common(x, rows, cols, dtype, shape2, axes2, norm, true, out, out_cap, out_shape2);
}

extern "C" __attribute__((visibility("default"))) int sfc_fftn(const void* x, int32_t ndim, const int64_t* in_shape, int dtype, const int64_t* shape,
                        const int64_t* axes, int32_t naxes, const char* norm, double* out, int64_t out_cap,
                        int64_t* out_shape) {
    return fftn_common(x, ndim, in_shape, dtype, shape, axes, naxes, norm, false, out, out_cap, out_shape, nullptr,
                       nullptr, nullptr);
}

extern "C" __attribute__((visibility("default"))) int sfc_ifftn(const void* x, int32_t ndim, const int64_t* in_shape, int dtype, const int64_t* shape,
                         const int64_t* axes, int32_t naxes, const char* norm, double* out, int64_t out_cap,
                         int64_t* out_shape) {
    return fftn_common(x, ndim, in_shape, dtype, shape, axes, naxes, norm, true, out, out_cap, out_shape, nullptr,
                       nullptr, nullptr);
}

// rfft2: full fft2(x, shape, None, None) then the first n_rows_out/2+1 ROWS (rfft.rs:212-232)
extern "C" __attribute__((visibility("default"))) int sfc_rfft2(const void* x, int64_t rows, int64_t cols, int dtype, const int64_t* shape2, double* out,
                         int64_t out_cap, int64_t* out_shape2) {
    int rc;
    if ((rc = require_device()) != SFC_OK) return rc;
    if (!dtype_ok(dtype)) return fail(SFC_ERR_VALUE, "unknown dtype");
    if (!x || rows <= 0 || cols <= 0) return fail(SFC_ERR_VALUE, "Input cannot be empty");
    const int64_t o0 = shape2 ? shape2[0] : rows, o1 = shape2 ? shape2[1] : cols;
    if (o0 <= 0 || o1 <= 0) return fail(SFC_ERR_VALUE, "FFT size must be positive");
    const int64_t r = o0 / 2 + 1;
    if (out_shape2) {
        out_shape2[0] = r;
        out_shape2[1] = o1;
    }
    if (!out || out_cap < r * o1) return fail(SFC_ERR_VALUE, "output buffer too small");
    void* d_res = nullptr;
    rc = run_c2c_host(x, {rows, cols}, dtype, {o0, o1}, {1, 0}, false, 1.0, &d_res);
    if (rc != SFC_OK) return rc;
    return download(out, d_res, (size_t)(r * o1) * 16);
}

// irfft2 (rfft.rs:274-355): zero-filled (rows_out, cols_out) spectrum, rows >= n_rows filled by
// conj(full[rows_out - i][(cols_out - j) % cols_out]) when that lies inside the input, ifft2,
// real part times (rows_out*cols_out)/(rows*cols).  The hard-coded 2x2 return is not reproduced.
extern "C" __attribute__((visibility("default"))) int sfc_irfft2(const void* x, int64_t rows, int64_t cols, int dtype, const int64_t* shape2, double* out,
                          int64_t out_cap, int64_t* out_shape2) {
    int rc;
    if ((rc = require_device()) != SFC_OK) return rc;
    if (!dtype_ok(dtype)) return fail(SFC_ERR_VALUE, "unknown dtype");
    if (!x || rows <= 0 || cols <= 0) return fail(SFC_ERR_VALUE, "Input cannot be empty");
    const int64_t o0 = shape2 ? shape2[0] : 2 * (rows - 1), o1 = shape2 ? shape2[1] : cols;
    if (o0 <= 0 || o1 <= 0) return fail(SFC_ERR_VALUE, "Input cannot be empty");
    if (rows > o0 || cols > o1)
        return fail(SFC_ERR_DIMENSION, "input extent exceeds the output shape (the reference indexes out of bounds here)");
    if (out_shape2) {
        out_shape2[0] = o0;
        out_shape2[1] = o1;
    }
    if (!out || out_cap < o0 * o1) return fail(SFC_ERR_VALUE, "output buffer too small");
    std::vector<double> full(2 * (size_t)(o0 * o1), 0.0);
    auto get = [&](int64_t i, int64_t j, double& re, double& im) {
        const int64_t k = i * cols + j;
        im = 0.0;
        switch (dtype) {
            case SFC_F32: re = ((const float*)x)[k]; break;
            case SFC_F64: re = ((const double*)x)[k]; break;
            case SFC_C64: re = ((const float*)x)[2 * k]; im = ((const float*)x)[2 * k + 1]; break;
            default: re = ((const double*)x)[2 * k]; im = ((const double*)x)[2 * k + 1]; break;
        }
    };
    for (int64_t i = 0; i < rows; ++i)
        for (int64_t j = 0; j < cols; ++j) get(i, j, full[2 * (i * o1 + j)], full[2 * (i * o1 + j) + 1]);
    for (int64_t i = rows; i < o0; ++i) {
        const int64_t si = o0 - i;
        for (int64_t j = 0; j < o1; ++j) {
            const int64_t sj = j == 0 ? 0 : o1 - j;
            if (si < rows && sj < cols) {
                full[2 * (i * o1 + j)] = full[2 * (si * o1 + sj)];
                full[2 * (i * o1 + j) + 1] = -full[2 * (si * o1 + sj) + 1];
            }
        }
    }
    const double scale = (1.0 / ((double)o0 * (double)o1)) * (((double)o0 * (double)o1) / ((double)rows * (double)cols));
    void* d_res = nullptr;
    rc = run_c2c_host(full.data(), {o0, o1}, SFC_C128, {o0, o1}, {1, 0}, true, scale, &d_res);
    if (rc != SFC_OK) return rc;
    std::vector<double> tmp(2 * (size_t)(o0 * o1));
    rc = download(tmp.data(), d_res, tmp.size() * 8);
    if (rc != SFC_OK) return rc;
    for (int64_t i = 0; i < o0 * o1; ++i) out[i] = tmp[2 * i];
    return SFC_OK;
}

// rfftn (rfft.rs:472-525): full fftn (with its all-dims norm quirk), then the last listed axis is
// cut to n/2+1 — only when `shape` is None (:508-511).
extern "C" __attribute__((visibility("default"))) int sfc_rfftn(const void* x, int32_t ndim, const int64_t* in_shape, int dtype, const int64_t* shape,
                         const int64_t* axes, int32_t naxes, const char* norm, double* out, int64_t out_cap,
                         int64_t* out_shape) {
    int rc;
    if ((rc = require_device()) != SFC_OK) return rc;
    if (!dtype_ok(dtype)) return fail(SFC_ERR_VALUE, "unknown dtype");
    if (ndim < 1 || ndim > SFC_MAX_DIMS) return fail(SFC_ERR_VALUE, "ndim must be in 1..8");
    if (!x || !in_shape) return fail(SFC_ERR_VALUE, "Input cannot be empty");
    std::vector<int64_t> ish(in_shape, in_shape + ndim);
    for (int64_t v : ish)
        if (v <= 0) return fail(SFC_ERR_VALUE, "Input cannot be empty");
    std::vector<int> ax;
    if (axes) {
        for (int i = 0; i < naxes; ++i) {
            if (axes[i] < 0 || axes[i] >= ndim) {
                char b[96];
                snprintf(b, sizeof b, "Axis %lld out of bounds for array of dimension %d", (long long)axes[i], ndim);
                return fail(SFC_ERR_VALUE, b);
            }
            ax.push_back((int)axes[i]);
        }
    } else {
        for (int i = 0; i < ndim; ++i) ax.push_back(i);
    }
    const int last_axis = ax.empty() ? ndim - 1 : ax.back();
    const NormMode nm = parse_norm_mode(norm, false);

    // fast path: real f64 input, shape None, at least one axis -> R2C plan (no full-size complex pass)
    if (!shape && dtype == SFC_F64 && !ax.empty()) {
        std::vector<int64_t> osh = ish;
        osh[last_axis] = ish[last_axis] / 2 + 1;
        if (out_shape)
            for (int i = 0; i < ndim; ++i) out_shape[i] = osh[i];
        if (!out || out_cap < vprod(osh)) return fail(SFC_ERR_VALUE, "output buffer too small");
        sfc_desc d;
        memset(&d, 0, sizeof d);
        d.ndim = ndim;
        for (int i = 0; i < ndim; ++i) d.shape[i] = ish[i];
        d.naxes = (int)ax.size();
        for (size_t i = 0; i < ax.size(); ++i) d.axes[i] = ax[i];
        d.kind = SFC_R2C;
        d.prec = SFC_PREC_F64;
        d.scale = norm_scale(nm, false, (double)vprod(ish));
        sfc_plan* h = nullptr;
        if ((rc = sfc_plan_create(&h, &d)) != SFC_OK) return rc;
        rc = sfc_exec_host(h, x, out);
        sfc_plan_destroy(h);
        return rc;
    }

    std::vector<int64_t> tsh;
    void* d_full = nullptr;
    rc = fftn_common(x, ndim, in_shape, dtype, shape, axes, naxes, norm, false, nullptr, 0, nullptr, &tsh, nullptr,
                     &d_full);
    if (rc != SFC_OK) return rc;
    std::vector<int64_t> osh = tsh;
    if (!shape) osh[last_axis] = tsh[last_axis] / 2 + 1;
    if (out_shape)
        for (int i = 0; i < ndim; ++i) out_shape[i] = osh[i];
    if (!out || out_cap < vprod(osh)) return fail(SFC_ERR_VALUE, "output buffer too small");
    if (osh == tsh) return download(out, d_full, (size_t)vprod(osh) * 16);
    void* d_crop = nullptr;
    if ((rc = g_ws.get(2, (size_t)vprod(osh) * 16, &d_crop)) != SFC_OK) return rc;
    CopyParams c;
    memset(&c, 0, sizeof c);
    c.ndim = ndim;
    for (int i = 0; i < ndim; ++i) {
        c.dst_shape[i] = osh[i];
        c.src_shape[i] = tsh[i];
    }
    c.src_complex = c.dst_complex = 1;
    c.src_f64 = c.dst_f64 = 1;
    c.scale = 1.0;
    c.total = vprod(osh);
    c.src = d_full;
    c.dst = d_crop;
    cudaError_t e = launch_nd_copy(c, g_ws.stream);
    if (e != cudaSuccess) return cuda_fail(e, "crop kernel");
    return download(out, d_crop, (size_t)vprod(osh) * 16);
}

// irfftn (rfft.rs:621-725)
extern "C" __attribute__((visibility("default"))) int sfc_irfftn(const void* x, int32_t ndim, const int64_t* in_shape, int dtype, const int64_t* shape,
                          int32_t nshape, const int64_t* axes, int32_t naxes, const char* norm, double* out,
                          int64_t out_cap, int64_t* out_shape) {
    int rc;
    if ((rc = require_device()) != SFC_OK) return rc;
    if (!dtype_ok(dtype)) return fail(SFC_ERR_VALUE, "unknown dtype");
    if (ndim < 1 || ndim > SFC_MAX_DIMS) return fail(SFC_ERR_VALUE, "ndim must be in 1..8");
    if (!x || !in_shape) return fail(SFC_ERR_VALUE, "Input cannot be empty");
    std::vector<int64_t> xsh(in_shape, in_shape + ndim);
    for (int64_t v : xsh)
        if (v <= 0) return fail(SFC_ERR_VALUE, "Input cannot be empty");
    std::vector<int> ax;
    if (axes) {
        for (int i = 0; i < naxes; ++i) {
            if (axes[i] < 0 || axes[i] >= ndim) {  // rfft.rs:642-648
                char b[112];
                snprintf(b, sizeof b, "Axis %lld is out of bounds for array of dimension %d", (long long)axes[i], ndim);
                return fail(SFC_ERR_DIMENSION, b);
            }
            ax.push_back((int)axes[i]);
        }
    } else {
        for (int i = 0; i < ndim; ++i) ax.push_back(i);
    }
    if ((int)ax.size() > SFC_MAX_DIMS) return fail(SFC_ERR_VALUE, "too many axes");
    std::vector<int64_t> osh;
    if (shape) {
        if (nshape != (int)ax.size() && !ax.empty() && nshape != ndim) {  // rfft.rs:659-672
            char b[200];
            snprintf(b, sizeof b,
                     "Shape must have the same number of dimensions as input or match the length of axes, got %d "
                     "expected %d or %d",
                     nshape, ndim, (int)ax.size());
            return fail(SFC_ERR_DIMENSION, b);
        }
        if (nshape == ndim) {
            osh.assign(shape, shape + ndim);
        } else if (nshape == (int)ax.size()) {
            osh = xsh;
            for (size_t i = 0; i < ax.size(); ++i) osh[ax[i]] = shape[i];
        } else {
            return fail(SFC_ERR_DIMENSION, "Shape has invalid dimensions");
        }
    } else {
        osh = xsh;
        const int last_axis = ax.empty() ? ndim - 1 : ax.back();
        osh[last_axis] = 2 * (osh[last_axis] - 1);  // rfft.rs:699-700
    }
    for (int64_t v : osh)
        if (v <= 0) return fail(SFC_ERR_VALUE, "Input cannot be empty");
    if (ax.empty()) return fail(SFC_ERR_VALUE, "irfftn needs at least one axis");  // reference panics on axes[0]
    for (int i = 0; i < ndim; ++i)
        if (xsh[i] > osh[i])
            return fail(SFC_ERR_DIMENSION,
                        "input extent exceeds the output shape (the reference indexes out of bounds here)");
    if (out_shape)
        for (int i = 0; i < ndim; ++i) out_shape[i] = osh[i];
    if (!out || out_cap < vprod(osh)) return fail(SFC_ERR_VALUE, "output buffer too small");

    // ifftn normalisation over the listed axes (algorithms.rs:876), default "backward"
    double total = 1.0;
    for (int a : ax) total *= (double)osh[a];
    const double scale = norm_scale(parse_norm_mode(norm, true), true, total);

    // widen to Complex64 on the device, then the C2R plan
    const int64_t xin = vprod(xsh);
    void *d_raw = nullptr, *d_x = nullptr, *d_o = nullptr;
    if ((rc = g_ws.get(0, (size_t)xin * dtype_bytes(dtype), &d_raw)) != SFC_OK) return rc;
    if ((rc = g_ws.get(1, (size_t)xin * 16, &d_x)) != SFC_OK) return rc;
    if ((rc = g_ws.get(2, (size_t)vprod(osh) * 8, &d_o)) != SFC_OK) return rc;
    cudaStream_t s = g_ws.stream;
    cudaError_t e = cudaMemcpyAsync(d_raw, x, (size_t)xin * dtype_bytes(dtype), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return cuda_fail(e, "H2D copy");
    const void* src = d_raw;
    if (dtype != SFC_C128) {
        CopyParams c;
        memset(&c, 0, sizeof c);
        c.ndim = ndim;
        for (int i = 0; i < ndim; ++i) c.dst_shape[i] = c.src_shape[i] = xsh[i];
        c.src_complex = dtype_is_complex(dtype);
        c.dst_complex = 1;
        c.src_f64 = dtype_is_f64(dtype);
        c.dst_f64 = 1;
        c.scale = 1.0;
        c.total = xin;
        c.src = d_raw;
        c.dst = d_x;
        e = launch_nd_copy(c, s);
        if (e != cudaSuccess) return cuda_fail(e, "convert kernel");
        src = d_x;
    }
    sfc_desc d;
    memset(&d, 0, sizeof d);
    d.ndim = ndim;
    for (int i = 0; i < ndim; ++i) {
        d.shape[i] = osh[i];
        d.in_shape[i] = xsh[i];
    }
    d.naxes = (int)ax.size();
    for (size_t i = 0; i < ax.size(); ++i) d.axes[i] = ax[i];
    d.kind = SFC_C2R;
    d.prec = SFC_PREC_F64;
    d.scale = scale;
    d.flags = SFC_DESC_CUSTOM_IN_SHAPE;
    PlanError perr{0, ""};
    std::shared_ptr<Plan> p = cached_plan(d, perr);
    if (!p) return fail(perr.code ? perr.code : SFC_ERR_PLAN, perr.msg);
    std::string es;
    rc = p->exec(src, d_o, s, es);
    if (rc != 0) return fail(rc, es);
    return download(out, d_o, (size_t)vprod(osh) * 8);
}

// strided_fft.rs:16-239 — one axis of an N-D array; ifft_strided scales by 1/len(axis)
extern "C" __attribute__((visibility("default"))) int sfc_fft_strided(const void* x, int32_t ndim, const int64_t* in_shape, int dtype, int64_t axis,
                               int inverse, double* out, int64_t out_cap) {
    int rc;
    if ((rc = require_device()) != SFC_OK) return rc;
    if (!dtype_ok(dtype)) return fail(SFC_ERR_VALUE, "unknown dtype");
    if (ndim < 1 || ndim > SFC_MAX_DIMS || !x || !in_shape) return fail(SFC_ERR_VALUE, "Input cannot be empty");
    if (axis < 0 || axis >= ndim) {  // strided_fft.rs:26-32
        char b[112];
        snprintf(b, sizeof b, "Axis %lld is out of bounds for array with %d dimensions", (long long)axis, ndim);
        return fail(SFC_ERR_VALUE, b);
    }
    std::vector<int64_t> sh(in_shape, in_shape + ndim);
    for (int64_t v : sh)
        if (v <= 0) return fail(SFC_ERR_VALUE, "Input cannot be empty");
    if (!out || out_cap < vprod(sh)) return fail(SFC_ERR_VALUE, "output buffer too small");
    void* d_res = nullptr;
    const double scale = inverse ? 1.0 / (double)sh[axis] : 1.0;
    rc = run_c2c_host(x, sh, dtype, sh, {(int)axis}, inverse != 0, scale, &d_res);
    if (rc != SFC_OK) return rc;
    return download(out, d_res, (size_t)vprod(sh) * 16);
}

// ----------------------------------------------------------- FftBackend trait

extern "C" __attribute__((visibility("default"))) int sfc_backend_fft_sized(const double* input, int64_t in_len, double* output, int64_t out_len,
                                     int64_t size) {
    if (in_len != size || out_len != size)  // backend.rs:96-100
        return fail(SFC_ERR_VALUE, "Input and output sizes must match the specified size");
    int64_t produced = 0;
    return sfc_fft(input, in_len, SFC_C128, size, output, out_len, &produced);
}

extern "C" __attribute__((visibility("default"))) int sfc_backend_ifft_sized(const double* input, int64_t in_len, double* output, int64_t out_len,
                                      int64_t size) {
    if (in_len != size || out_len != size)  // backend.rs:129-133
        return fail(SFC_ERR_VALUE, "Input and output sizes must match the specified size");
    int64_t produced = 0;
    return sfc_ifft(input, in_len, SFC_C128, size, output, out_len, &produced);  // 1/n as backend.rs:149-152
}

extern "C" __attribute__((visibility("default"))) int sfc_backend_fft(const double* input, int64_t in_len, double* output, int64_t out_len) {
    return sfc_backend_fft_sized(input, in_len, output, out_len, in_len);
}

extern "C" __attribute__((visibility("default"))) int sfc_backend_ifft(const double* input, int64_t in_len, double* output, int64_t out_len) {
    return sfc_backend_ifft_sized(input, in_len, output, out_len, in_len);
}

extern "C" __attribute__((visibility("default"))) int sfc_backend_supports_feature(const char* feature) {
    if (!feature) return 0;
    static const char* k[] = {"1d_fft", "2d_fft", "nd_fft", "cached_plans", "gpu_acceleration", "batched", "f32"};
    for (const char* f : k)
        if (!strcmp(f, feature)) return 1;
    return 0;
}

extern "C" __attribute__((visibility("default"))) const char* sfc_backend_name(void) { return "cuda_fft"; }
extern "C" __attribute__((visibility("default"))) const char* sfc_backend_description(void) {
    return "B200-native CUDA FFT (sm_100a Stockham tile kernels, four-step, Bluestein)";
}

// ------------------------------- ParallelExecutor::execute_batch (planning_parallel.rs:316-405)

extern "C" __attribute__((visibility("default"))) int sfc_execute_batch(const double* inputs, double* outputs, int64_t count, int64_t size, int inverse) {
    int rc;
    if ((rc = require_device()) != SFC_OK) return rc;
    if (!inputs || !outputs || count <= 0 || size <= 0) return fail(SFC_ERR_VALUE, "Input cannot be empty");
    sfc_desc d;
    memset(&d, 0, sizeof d);
    d.ndim = 2;
    d.shape[0] = count;
    d.shape[1] = size;
    d.naxes = 1;
    d.axes[0] = 1;
    d.kind = SFC_C2C;
    d.prec = SFC_PREC_F64;
    d.direction = inverse ? SFC_INVERSE : SFC_FORWARD;
    d.scale = 1.0;  // FftPlanExecutor::execute is unnormalised in both directions (planning.rs:501-550)
    {
        bool handled = false;  // sfc_set_num_gpus(P > 1): contiguous batch split over the GPUs of this process
        rc = multi_batch_host(d, inputs, outputs, &handled);
        if (rc != SFC_OK || handled) return rc;
    }
    sfc_plan* h = nullptr;
    if ((rc = sfc_plan_create(&h, &d)) != SFC_OK) return rc;
    rc = sfc_exec_host(h, inputs, outputs);
    sfc_plan_destroy(h);
    return rc;
}

// ---------------------------------------------------- batched real transforms (f32 / f64 compute)

static int real_batch(const void* x, int64_t batch, int64_t n, int prec, void* out, bool inverse) {
    int rc;
    if ((rc = require_device()) != SFC_OK) return rc;
    if (!x || !out || batch <= 0 || n <= 0) return fail(SFC_ERR_VALUE, "Input cannot be empty");
    if (prec != SFC_PREC_F32 && prec != SFC_PREC_F64) return fail(SFC_ERR_VALUE, "unknown precision");
    sfc_desc d;
    memset(&d, 0, sizeof d);
    d.ndim = 2;
    d.shape[0] = batch;
    d.shape[1] = n;
    d.naxes = 1;
    d.axes[0] = 1;
    d.kind = inverse ? SFC_C2R : SFC_R2C;
    d.prec = prec;
    d.scale = inverse ? 1.0 / (double)n : 1.0;
    sfc_plan* h = nullptr;
    if ((rc = sfc_plan_create(&h, &d)) != SFC_OK) return rc;
    rc = sfc_exec_host(h, x, out);
    sfc_plan_destroy(h);
    return rc;
}

extern "C" __attribute__((visibility("default"))) int sfc_rfft_batch(const void* x, int64_t batch, int64_t n, int prec, void* out) {
    return real_batch(x, batch, n, prec, out, false);
}

extern "C" __attribute__((visibility("default"))) int sfc_irfft_batch(const void* x, int64_t batch, int64_t n, int prec, void* out) {
    return real_batch(x, batch, n, prec, out, true);
}
