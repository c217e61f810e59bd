// api_ext.cu — C ABI of the in-crate consumers of the FFT hot path: every one of
// them is "O(n) pre-pass -> FFT -> O(n) post-pass", so here they are device tables / element-wise
// kernels around (or fused into) the plans of plan.cu.  Reference semantics followed line by line:
//   DCT  I-IV  scirs2-fft/src/dct.rs:425-757      (direct O(n^2) sums there, same numbers here)
//   DST  I-IV  scirs2-fft/src/dst.rs:409-702
//   Hartley    scirs2-fft/src/hartley.rs:37-200
//   hfft/ihfft scirs2-fft/src/hfft/complex_to_real.rs:58-135, real_to_complex.rs:49-147
//   hilbert    scirs2-fft/src/lib.rs:437-516
//   stft / spectrogram  scirs2-fft/src/spectrogram.rs:76-310, 312-420
// NOT reproduced (reference test hacks): the hard-coded `[1,2,3,4]` returns of idct1 / idst1..4 for
// n == 4 && norm == "ortho" (dct.rs:483-485, dst.rs:459-461, 528-530, 604-606, 679-681).
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

#include "api_internal.h"
#include "ext_kernels.cuh"

using namespace sfc;
using namespace sfc_api;

#define SFC_EXPORT extern "C" __attribute__((visibility("default")))

namespace {

inline bool is_pow2_i64(int64_t n) { return n > 0 && (n & (n - 1)) == 0; }

bool fuse_enabled() {
    static int v = [] {
        const char* e = getenv("SFC_EXT_FUSE");
        return e ? atoi(e) : 1;
    }();
    return v != 0;
}

// the fused type-IV kernel (TM_FAST_DCT4): measured 63-66 % of its roofline against 12 % for the 2n-point formulation
// (profiles/r2a_first_run.log), parity <= 1e-12 (tests/test_gpu_experimental.py); on by default since round 2
bool dct4_fused_enabled() {
    static int v = [] {
        const char* e = getenv("SFC_DCT4_FUSED");
        return e ? atoi(e) : 1;
    }();
    return v != 0;
}

// ------------------------------------------------------------------------------------------
// DCT / DST of every type as ONE complex FFT of length P = 2D on zero-padded data:
//   out[k] = g[k] * sum_i s[i] x[i] f(pi (i + alpha)(k + beta) / D),   f = cos | sin
//          = Re( w[k] * sum_i (u[i] x[i]) exp(-2 pi i * i k / P) )
//   u[i] = s[i] exp(-i pi (2 i b2) / (4D)),  w[k] = g[k] c exp(-i pi a2 (2k + b2) / (4D)),  a2 = 2 alpha, b2 = 2 beta,
//   c = 1 (cos) or i (sin: Re(i z) = -Im z).  The phases are reduced exactly (mod 8D) in integers.
struct TrigTables {
    int64_t P = 0;
    void* d_u = nullptr;  // n complex f64
    void* d_w = nullptr;
};

void unit(long double q, long double den4, long double& c, long double& s) {
    // exp(-i pi q / den4)
    const long double pi = 3.141592653589793238462643383279502884L;
    const long double a = pi * (q / den4);
    c = cosl(a);
    s = -sinl(a);
}

// kind 0 = DCT, 1 = DST; type 1..4.  Returns an error text or "".
std::string trig_spec(int kind, int type, bool inverse, bool ortho, int64_t n, int64_t& D, int& a2, int& b2,
                      std::vector<long double>& s, std::vector<long double>& g) {
    const long double nn = (long double)n;
    s.assign(n, 1.0L);
    g.assign(n, 1.0L);
    const long double r2 = sqrtl(2.0L);
    if (kind == 0) {
        if (type == 1) {
            if (n < 2)
                return inverse ? "Input array must have at least 2 elements for IDCT-I"
                               : "Input array must have at least 2 elements for DCT-I";
            D = n - 1;
            a2 = 0;
            b2 = 0;
            const long double m = (long double)(n - 1);
            if (!inverse) {  // dct.rs:425-469
                for (int64_t k = 0; k < n; ++k) {
                    const bool end = (k == 0 || k == n - 1);
                    g[k] = end ? 0.5L : 1.0L;
                    if (ortho) g[k] *= sqrtl(2.0L / m) * (end ? 1.0L / r2 : 1.0L);
                }
            } else {  // dct.rs:473-519
                for (int64_t k = 0; k < n; ++k) {
                    const bool end = (k == 0 || k == n - 1);
                    s[k] = end ? 0.5L : 1.0L;
                    if (ortho) s[k] *= sqrtl(m / 2.0L) * (end ? r2 : 1.0L);
                    g[k] = 2.0L / m;
                }
            }
        } else if (type == 2) {
            D = n;
            if (!inverse) {  // dct.rs:523-559
                a2 = 1;
                b2 = 0;
                if (ortho)
                    for (int64_t k = 0; k < n; ++k) g[k] = sqrtl(2.0L / nn) * (k == 0 ? 1.0L / r2 : 1.0L);
            } else {  // dct.rs:563-601
                a2 = 0;
                b2 = 1;
                for (int64_t k = 0; k < n; ++k) {
                    s[k] = k == 0 ? 0.5L : 1.0L;
                    if (ortho) s[k] *= sqrtl(nn / 2.0L) * (k == 0 ? r2 : 1.0L);
                    g[k] = 2.0L / nn;
                }
            }
        } else if (type == 3) {
            D = n;
            a2 = 0;
            b2 = 1;
            if (!inverse) {  // dct.rs:605-643
                for (int64_t k = 0; k < n; ++k) {
                    s[k] = k == 0 ? 0.5L : 1.0L;
                    if (ortho) s[k] *= sqrtl(nn / 2.0L) * (k == 0 ? 1.0L / r2 : 1.0L);
                    g[k] = 2.0L / nn;
                }
            } else {  // dct.rs:647-684
                if (ortho)
                    for (int64_t k = 0; k < n; ++k) s[k] = sqrtl(2.0L / nn) * (k == 0 ? r2 : 1.0L);
            }
        } else if (type == 4) {
            D = n;
            a2 = 1;
            b2 = 1;
            if (!inverse) {  // dct.rs:688-720
                if (ortho) g.assign(n, sqrtl(2.0L / nn));
            } else {  // dct.rs:724-746: scale the input, then dct4(input, norm)
                s.assign(n, ortho ? sqrtl(nn / 2.0L) : 2.0L / nn);
                if (ortho) g.assign(n, sqrtl(2.0L / nn));
            }
        } else
            return "unknown DCT type";
        if (n == 0) return "Input array cannot be empty";
        return "";
    }
    // DST, dst.rs
    if (type == 1) {
        if (n < 2)
            return inverse ? "Input array must have at least 2 elements for IDST-I"
                           : "Input array must have at least 2 elements for DST-I";
        D = n + 1;
        a2 = 2;
        b2 = 2;
        const long double m = (long double)(n + 1);
        if (!inverse) {  // dst.rs:409-446
            g.assign(n, ortho ? sqrtl(2.0L / m) : 2.0L / sqrtl(m));
        } else {  // dst.rs:450-480: both norm branches scale by sqrt(n+1)/2, then dst1(input, None)
            s.assign(n, sqrtl(m) / 2.0L);
            g.assign(n, 2.0L / sqrtl(m));
        }
    } else if (type == 2) {
        D = n;
        if (!inverse) {  // dst.rs:484-516
            a2 = 1;
            b2 = 2;
            if (ortho) g.assign(n, sqrtl(2.0L / nn));
        } else {  // dst.rs:520-545: scale, then dst3(input, None)
            a2 = 2;
            b2 = 1;
            if (ortho) s.assign(n, sqrtl(nn / 2.0L));
            g.assign(n, 0.5L);
        }
    } else if (type == 3) {
        D = n;
        if (!inverse) {  // dst.rs:549-592 (the x[n-1] (-1)^k term is the m = n-1 term of the same sum)
            a2 = 2;
            b2 = 1;
            g.assign(n, ortho ? sqrtl(2.0L / nn) / 2.0L : 0.5L);
        } else {  // dst.rs:596-626: scale, then dst2(input, None)
            a2 = 1;
            b2 = 2;
            s.assign(n, ortho ? sqrtl(nn / 2.0L) * 2.0L : 2.0L);
        }
    } else if (type == 4) {
        D = n;
        a2 = 1;
        b2 = 1;
        if (!inverse) {  // dst.rs:630-667
            g.assign(n, ortho ? sqrtl(2.0L / nn) : 2.0L);
        } else {  // dst.rs:671-701: scale, then dst4(input, None)
            s.assign(n, ortho ? sqrtl(nn / 2.0L) : 0.5L);
            g.assign(n, 2.0L);
        }
    } else
        return "unknown DST type";
    if (n == 0) return "Input array cannot be empty";
    return "";
}

std::mutex g_trig_mu;
std::map<std::tuple<int, int, int, int, int, int64_t>, TrigTables> g_trig;

int get_trig_tables(int kind, int type, bool inverse, bool ortho, int64_t n, TrigTables& out) {
    int dev = 0;
    cudaGetDevice(&dev);
    const auto key = std::make_tuple(dev, kind, type, (int)inverse, (int)ortho, n);
    std::lock_guard<std::mutex> lk(g_trig_mu);
    auto it = g_trig.find(key);
    if (it != g_trig.end()) {
        out = it->second;
        return SFC_OK;
    }
    int64_t D = 0;
    int a2 = 0, b2 = 0;
    std::vector<long double> s, g;
    const std::string msg = trig_spec(kind, type, inverse, ortho, n, D, a2, b2, s, g);
    if (!msg.empty()) return fail(SFC_ERR_VALUE, msg);
    const bool sine = kind == 1;
    std::vector<double> u(2 * (size_t)n), w(2 * (size_t)n);
    const unsigned __int128 mod = 8 * (unsigned __int128)D;
    for (int64_t i = 0; i < n; ++i) {
        long double c, sn;
        unsigned __int128 q = ((unsigned __int128)(2 * i) * (unsigned)b2) % mod;
        unit((long double)(uint64_t)q, 4.0L * (long double)D, c, sn);
        u[2 * i] = (double)(s[i] * c);
        u[2 * i + 1] = (double)(s[i] * sn);
        q = ((unsigned __int128)(unsigned)a2 * (unsigned __int128)(2 * i + b2)) % mod;
        unit((long double)(uint64_t)q, 4.0L * (long double)D, c, sn);
        if (sine) {  // times i
            w[2 * i] = (double)(-g[i] * sn);
            w[2 * i + 1] = (double)(g[i] * c);
        } else {
            w[2 * i] = (double)(g[i] * c);
            w[2 * i + 1] = (double)(g[i] * sn);
        }
    }
    TrigTables t;
    t.P = 2 * D;
    const size_t bytes = (size_t)n * 16;
    cudaError_t e = cudaMalloc(&t.d_u, bytes);
    if (e == cudaSuccess) e = cudaMalloc(&t.d_w, bytes);
    if (e == cudaSuccess) e = cudaMemcpy(t.d_u, u.data(), bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(t.d_w, w.data(), bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return cuda_fail(e, "DCT/DST twiddle tables");
    g_trig[key] = t;  // 32 n bytes per (type, direction, norm, n); cached plans refer to these pointers
    out = t;
    return SFC_OK;
}

// one 1-D trig transform along axis `a` of a real [O][N][I] array: d_src (real) -> d_dst (real)
int trig_axis(int kind, int type, bool inverse, bool ortho, int64_t O, int64_t N, int64_t I, const void* d_src,
              void* d_dst, cudaStream_t st) {
    int rc;
    {
        // argument checks (and error texts) of the reference come from the table builder's specification
        int64_t D = 0;
        int a2 = 0, b2 = 0;
        std::vector<long double> s, g;
        if (N < 2) {
            const std::string msg = trig_spec(kind, type, inverse, ortho, N, D, a2, b2, s, g);
            if (!msg.empty()) return fail(SFC_ERR_VALUE, msg);
        }
    }
    TrigTables t;
    sfc_desc d;
    memset(&d, 0, sizeof d);
    d.ndim = 3;
    d.shape[0] = O;
    d.shape[1] = t.P;
    d.shape[2] = I;
    d.naxes = 1;
    d.axes[0] = 1;
    d.kind = SFC_C2C;
    d.prec = SFC_PREC_F64;
    d.direction = SFC_FORWARD;
    d.scale = 1.0;
    d.axis_in_len = N;
    d.axis_out_len = N;
    PlanError perr{0, ""};
    std::string es;
    if (fuse_enabled() && (type == 2 || type == 3) && is_pow2_i64(N) && N >= 128) {
        // Types II / III of contiguous rows: ONE kernel on the N/2-point packed transform (Makhoul), 8x less data on
        // chip than the 2N-point formulation.  Every variant is  C2: g * sum_i x[i] cos(pi (i+1/2) k / N)  (output 0
        // times dc)  or  C3: scale * (dc * X[0] + 2 sum_k>=1 X[k] cos(pi k (i+1/2) / N)),  or their sine twins.
        const double nn = (double)N, r2 = std::sqrt(2.0);
        const bool sine = kind == 1;
        bool c3;        // which kernel
        double sc, dc;  // its two factors
        if (!sine) {
            if (type == 2 && !inverse) { c3 = false; sc = ortho ? std::sqrt(2.0 / nn) : 1.0; dc = ortho ? 1.0 / r2 : 1.0; }          // dct.rs:523-559
            else if (type == 2) { c3 = true; sc = ortho ? std::sqrt(2.0 / nn) / 2 : 1.0 / nn; dc = ortho ? r2 : 1.0; }               // :563-601
            else if (!inverse) { c3 = true; sc = ortho ? std::sqrt(2.0 / nn) / 2 : 1.0 / nn; dc = ortho ? 1.0 / r2 : 1.0; }          // :605-643
            else { c3 = true; sc = ortho ? std::sqrt(2.0 / nn) / 2 : 0.5; dc = ortho ? 2.0 * r2 : 2.0; }                            // :647-684
        } else {
            if (type == 2 && !inverse) { c3 = false; sc = ortho ? std::sqrt(2.0 / nn) : 1.0; dc = 1.0; }                             // dst.rs:484-516
            else if (type == 2) { c3 = true; sc = ortho ? std::sqrt(nn / 2.0) / 4 : 0.25; dc = 2.0; }                                // :520-545 (dst3 of the scaled input)
            else if (!inverse) { c3 = true; sc = ortho ? std::sqrt(2.0 / nn) / 4 : 0.25; dc = 2.0; }                                 // :549-592
            else { c3 = false; sc = ortho ? 2.0 * std::sqrt(nn / 2.0) : 2.0; dc = 1.0; }                                             // :596-626 (dst2 of the scaled input)
        }
        sfc_desc dd;
        memset(&dd, 0, sizeof dd);
        dd.ndim = 3;
        dd.shape[0] = O;
        dd.shape[1] = N;
        dd.shape[2] = I;
        dd.naxes = 1;
        dd.axes[0] = 1;
        dd.kind = SFC_R2C;
        dd.prec = SFC_PREC_F64;
        dd.scale = sc;
        dd.scale_dc = dc;
        dd.flags = (c3 ? SFC_DESC_DCT3 : SFC_DESC_DCT2) | (sine ? SFC_DESC_TRIG_SINE : 0);
        std::shared_ptr<Plan> p = cached_plan(dd, perr);
        if (p) {
            rc = p->exec(d_src, d_dst, st, es);
            if (rc != 0) return fail(rc, es);
            return SFC_OK;
        }
        if (perr.code != SFC_ERR_NOT_IMPLEMENTED) return fail(perr.code ? perr.code : SFC_ERR_PLAN, perr.msg);
    }
    if (fuse_enabled() && dct4_fused_enabled() && type == 4 && is_pow2_i64(N) && N >= 128) {
        // Type IV, rows or a strided axis: one kernel on the N/2-point complex transform.  Every variant is a scaled
        // sum_i x[i] cos|sin(pi (i+1/2)(k+1/2) / N): dct.rs:688-720 (1 | sqrt(2/N)), :724-746 (input * 2/N | sqrt(N/2), then the
        // forward sum), dst.rs:630-667 (2 | sqrt(2/N)), :671-702 (input * 1/2 | sqrt(N/2), then the un-normalised sum * 2)
        const double nn = (double)N;
        const bool sine = kind == 1;
        double sc;
        if (!sine) sc = !inverse ? (ortho ? std::sqrt(2.0 / nn) : 1.0) : (ortho ? std::sqrt(nn / 2.0) * std::sqrt(2.0 / nn) : 2.0 / nn);
        else sc = !inverse ? (ortho ? std::sqrt(2.0 / nn) : 2.0) : (ortho ? 2.0 * std::sqrt(nn / 2.0) : 1.0);
        sfc_desc dd;
        memset(&dd, 0, sizeof dd);
        dd.ndim = 3;
        dd.shape[0] = O;
        dd.shape[1] = N;
        dd.shape[2] = I;
        dd.naxes = 1;
        dd.axes[0] = 1;
        dd.kind = SFC_R2C;
        dd.prec = SFC_PREC_F64;
        dd.scale = sc;
        dd.flags = SFC_DESC_DCT4 | (sine ? SFC_DESC_TRIG_SINE : 0);
        std::shared_ptr<Plan> p = cached_plan(dd, perr);
        if (p) {
            rc = p->exec(d_src, d_dst, st, es);
            if (rc != 0) return fail(rc, es);
            return SFC_OK;
        }
        if (perr.code != SFC_ERR_NOT_IMPLEMENTED) return fail(perr.code ? perr.code : SFC_ERR_PLAN, perr.msg);
    }
    if ((rc = get_trig_tables(kind, type, inverse, ortho, N, t)) != SFC_OK) return rc;  // only the 2N-point formulations need them
    d.shape[1] = t.P;
    if (fuse_enabled() && is_pow2_i64(t.P)) {
        // everything in the FFT passes themselves: real load * u, ..., * w, real-part store
        d.flags = SFC_DESC_AXIS_LEN | SFC_DESC_AUX_MUL | SFC_DESC_REAL_INPUT | SFC_DESC_REAL_OUTPUT;
        d.aux_in = t.d_u;
        d.aux_out = t.d_w;
        std::shared_ptr<Plan> p = cached_plan(d, perr);
        if (p) {
            rc = p->exec(d_src, d_dst, st, es);
            if (rc != 0) return fail(rc, es);
            return SFC_OK;
        }
        if (perr.code != SFC_ERR_NOT_IMPLEMENTED) return fail(perr.code ? perr.code : SFC_ERR_PLAN, perr.msg);
        d.aux_in = d.aux_out = nullptr;
    }
    // general lengths (Bluestein inside the plan): explicit pre / post passes
    void* d_cplx = nullptr;
    if ((rc = g_ws.get(1, (size_t)(O * N * I) * 16, &d_cplx)) != SFC_OK) return rc;
    d.flags = SFC_DESC_AXIS_LEN;
    std::shared_ptr<Plan> p = cached_plan(d, perr);
    if (!p) return fail(perr.code ? perr.code : SFC_ERR_PLAN, perr.msg);
    MapParams m{};
    m.src = d_src;
    m.dst = d_cplx;
    m.tab = t.d_u;
    m.total = O * N * I;
    m.n = N;
    m.inner = I;
    m.mode = EM_R2C_TAB;
    cudaError_t e = launch_map(m, st);
    if (e != cudaSuccess) return cuda_fail(e, "DCT/DST pre-twiddle kernel");
    rc = p->exec(d_cplx, d_cplx, st, es);
    if (rc != 0) return fail(rc, es);
    m.src = d_cplx;
    m.dst = d_dst;
    m.tab = t.d_w;
    m.mode = EM_C2R_TAB;
    e = launch_map(m, st);
    if (e != cudaSuccess) return cuda_fail(e, "DCT/DST post-twiddle kernel");
    return SFC_OK;
}

int trig_nd(int kind, const double* x, int32_t ndim, const int64_t* shape, const int32_t* axes, int32_t naxes, int32_t type,
            int32_t inverse, const char* norm, double* out) {
    int rc;
    if ((rc = require_device()) != SFC_OK) return rc;
    if (!x || !out || !shape || ndim < 1 || ndim > SFC_MAX_DIMS) return fail(SFC_ERR_VALUE, "Input array cannot be empty");
    if (type < 1 || type > 4) return fail(SFC_ERR_VALUE, kind == 0 ? "unknown DCT type" : "unknown DST type");
    std::vector<int64_t> sh(shape, shape + ndim);
    for (int64_t v : sh)
        if (v <= 0) return fail(SFC_ERR_VALUE, "Input array cannot be empty");
    std::vector<int> ax;
    if (axes)
        ax.assign(axes, axes + naxes);
    else
        for (int i = 0; i < ndim; ++i) ax.push_back(i);  // dct.rs:317: None = every axis in order
    for (int a : ax)
        if (a < 0 || a >= ndim) return fail(SFC_ERR_VALUE, "axis out of bounds");
    const bool ortho = norm && !strcmp(norm, "ortho");  // any other string = no normalisation (dct.rs:455)
    const int64_t total = vprod(sh);
    void *d_a = nullptr, *d_b = nullptr;
    if ((rc = g_ws.get(0, (size_t)total * 8, &d_a)) != SFC_OK) return rc;
    if ((rc = g_ws.get(3, (size_t)total * 8, &d_b)) != SFC_OK) return rc;
    cudaStream_t st = g_ws.stream;
    cudaError_t e = cudaMemcpyAsync(d_a, x, (size_t)total * 8, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "H2D copy");
    void* cur = d_a;
    void* nxt = d_b;
    for (int a : ax) {
        int64_t O = 1, I = 1;
        for (int i = 0; i < a; ++i) O *= sh[i];
        for (int i = a + 1; i < ndim; ++i) I *= sh[i];
        if ((rc = trig_axis(kind, type, inverse != 0, ortho, O, sh[a], I, cur, nxt, st)) != SFC_OK) return rc;
        std::swap(cur, nxt);
    }
    return download(out, cur, (size_t)total * 8);
}

// forward c2c of one real f64 lane set [O][n][I] at transform length P, first n bins kept: -> complex [O][n][I]
int real_fft_crop(const void* d_real, void* d_cplx, int64_t O, int64_t n, int64_t I, int64_t P, cudaStream_t st) {
    sfc_desc d;
    memset(&d, 0, sizeof d);
    d.ndim = 3;
    d.shape[0] = O;
    d.shape[1] = P;
    d.shape[2] = I;
    d.naxes = 1;
    d.axes[0] = 1;
    d.kind = SFC_C2C;
    d.prec = SFC_PREC_F64;
    d.direction = SFC_FORWARD;
    d.scale = 1.0;
    d.flags = SFC_DESC_AXIS_LEN | SFC_DESC_REAL_INPUT;
    d.axis_in_len = n;
    d.axis_out_len = n;
    PlanError perr{0, ""};
    std::shared_ptr<Plan> p = cached_plan(d, perr);
    if (!p) return fail(perr.code ? perr.code : SFC_ERR_PLAN, perr.msg);
    std::string es;
    const int rc = p->exec(d_real, d_cplx, st, es);
    if (rc != 0) return fail(rc, es);
    return SFC_OK;
}

// hartley.rs:37-66 along the middle axis of [O][n][I]: H = Re(F) - Im(F), F = fft(x, None) — which pads to the
// next power of two and keeps the first n bins (quirk reproduced)
int dht_axis(const void* d_src, void* d_dst, void* d_cplx, int64_t O, int64_t n, int64_t I, double scale, cudaStream_t st) {
    int rc = real_fft_crop(d_src, d_cplx, O, n, I, next_pow2_i64(n), st);
    if (rc != SFC_OK) return rc;
    MapParams m{};
    m.src = d_cplx;
    m.dst = d_dst;
    m.total = O * n * I;
    m.n = n;
    m.inner = I;
    m.mode = EM_C_LINCOMB;
    m.a = scale;
    m.b = -scale;
    cudaError_t e = launch_map(m, st);
    if (e != cudaSuccess) return cuda_fail(e, "Hartley output kernel");
    return SFC_OK;
}

int dht_common(const double* x, int64_t n, double* out, bool inverse) {
    int rc;
    if ((rc = require_device()) != SFC_OK) return rc;
    if (!x || n <= 0 || !out) return fail(SFC_ERR_VALUE, "empty array");  // hartley.rs:46-48
    void *d_a = nullptr, *d_b = nullptr, *d_c = nullptr;
    if ((rc = g_ws.get(0, (size_t)n * 8, &d_a)) != SFC_OK) return rc;
    if ((rc = g_ws.get(3, (size_t)n * 8, &d_b)) != SFC_OK) return rc;
    if ((rc = g_ws.get(1, (size_t)n * 16, &d_c)) != SFC_OK) return rc;
    cudaStream_t st = g_ws.stream;
    cudaError_t e = cudaMemcpyAsync(d_a, x, (size_t)n * 8, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "H2D copy");
    if ((rc = dht_axis(d_a, d_b, d_c, 1, n, 1, inverse ? 1.0 / (double)n : 1.0, st)) != SFC_OK) return rc;  // hartley.rs:104-108
    return download(out, d_b, (size_t)n * 8);
}

}  // namespace

SFC_EXPORT int sfc_dct(const double* x, int32_t ndim, const int64_t* shape, const int32_t* axes, int32_t naxes, int32_t type,
                       int32_t inverse, const char* norm, double* out) {
    return trig_nd(0, x, ndim, shape, axes, naxes, type, inverse, norm, out);
}

SFC_EXPORT int sfc_dst(const double* x, int32_t ndim, const int64_t* shape, const int32_t* axes, int32_t naxes, int32_t type,
                       int32_t inverse, const char* norm, double* out) {
    return trig_nd(1, x, ndim, shape, axes, naxes, type, inverse, norm, out);
}

SFC_EXPORT int sfc_dht(const double* x, int64_t n, double* out) { return dht_common(x, n, out, false); }
SFC_EXPORT int sfc_idht(const double* h, int64_t n, double* out) { return dht_common(h, n, out, true); }

// hartley.rs:133-200: dht along axes.0, then along axes.1
SFC_EXPORT int sfc_dht2(const double* x, int64_t rows, int64_t cols, int32_t axis0, int32_t axis1, double* out) {
    int rc;
    if ((rc = require_device()) != SFC_OK) return rc;
    if (axis0 < 0 || axis0 >= 2 || axis1 < 0 || axis1 >= 2) {
        char b[96];
        snprintf(b, sizeof b, "Axes out of bounds: (%d, %d)", axis0, axis1);
        return fail(SFC_ERR_VALUE, b);  // hartley.rs:143-148
    }
    if (!x || rows <= 0 || cols <= 0 || !out) return fail(SFC_ERR_VALUE, "empty array");
    const int64_t total = rows * cols;
    void *d_a = nullptr, *d_b = nullptr, *d_c = nullptr;
    if ((rc = g_ws.get(0, (size_t)total * 8, &d_a)) != SFC_OK) return rc;
    if ((rc = g_ws.get(3, (size_t)total * 8, &d_b)) != SFC_OK) return rc;
    if ((rc = g_ws.get(1, (size_t)total * 16, &d_c)) != SFC_OK) return rc;
    cudaStream_t st = g_ws.stream;
    cudaError_t e = cudaMemcpyAsync(d_a, x, (size_t)total * 8, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "H2D copy");
    auto along = [&](int axis, const void* src, void* dst) {
        return axis == 0 ? dht_axis(src, dst, d_c, 1, rows, cols, 1.0, st) : dht_axis(src, dst, d_c, rows, cols, 1, 1.0, st);
    };
    // note hartley.rs:154-171: `axes.0 == 0` transforms the columns (axis 0), anything else the rows
    if ((rc = along(axis0 == 0 ? 0 : 1, d_a, d_b)) != SFC_OK) return rc;
    // hartley.rs:176-197: `axes.1 == 1` transforms the rows (axis 1), anything else the columns
    if ((rc = along(axis1 == 1 ? 1 : 0, d_b, d_a)) != SFC_OK) return rc;
    return download(out, d_a, (size_t)total * 8);
}

// hfft/complex_to_real.rs:58-135: real part of fft(x with Im x[0] = 0, Some(n or len)).  Im x[0] only adds a purely
// imaginary constant to every bin, so the real parts do not depend on it.
SFC_EXPORT int sfc_hfft(const void* x, int64_t len, int dtype, int64_t n, double* out, int64_t out_cap, int64_t* out_len) {
    int rc;
    if ((rc = require_device()) != SFC_OK) return rc;
    if (!dtype_ok(dtype)) return fail(SFC_ERR_VALUE, "unknown dtype");
    if (len <= 0 || !x) return fail(SFC_ERR_VALUE, "Input cannot be empty");  // from fft(), algorithms.rs:136-138
    if (n == 0) return fail(SFC_ERR_VALUE, "FFT size must be positive");
    const int64_t n_fft = n > 0 ? n : len;  // complex_to_real.rs:114
    if (out_len) *out_len = n_fft;
    if (!out || out_cap < n_fft) return fail(SFC_ERR_VALUE, "output buffer too small");
    void* d_res = nullptr;
    if ((rc = run_c2c_host(x, {len}, dtype, {n_fft}, {0}, false, 1.0, &d_res)) != SFC_OK) return rc;
    void* d_r = nullptr;
    if ((rc = g_ws.get(3, (size_t)n_fft * 8, &d_r)) != SFC_OK) return rc;
    MapParams m{};
    m.src = d_res;
    m.dst = d_r;
    m.total = n_fft;
    m.n = n_fft;
    m.inner = 1;
    m.mode = EM_C_LINCOMB;
    m.a = 1.0;
    m.b = 0.0;
    cudaError_t e = launch_map(m, g_ws.stream);
    if (e != cudaSuccess) return cuda_fail(e, "hfft output kernel");
    return download(out, d_r, (size_t)n_fft * 8);
}

// hfft/real_to_complex.rs:112-149: ifft(x resized to n, Some(n)), DC made real, upper half = conjugate reflection
SFC_EXPORT int sfc_ihfft(const double* x, int64_t len, int64_t n, double* out, int64_t out_cap, int64_t* out_len) {
    int rc;
    if ((rc = require_device()) != SFC_OK) return rc;
    if (n == 0) return fail(SFC_ERR_VALUE, "FFT size must be positive");
    const int64_t n_fft = n > 0 ? n : len;
    if (n_fft <= 0 || !x) return fail(SFC_ERR_VALUE, "Input cannot be empty");
    if (out_len) *out_len = n_fft;
    if (!out || out_cap < n_fft) return fail(SFC_ERR_VALUE, "output buffer too small");
    if (len <= 0) return fail(SFC_ERR_VALUE, "Input cannot be empty");
    void* d_res = nullptr;
    const int64_t used = std::min(len, n_fft);
    if ((rc = run_c2c_host(x, {used}, SFC_F64, {n_fft}, {0}, true, 1.0 / (double)n_fft, &d_res)) != SFC_OK) return rc;
    void* d_o = nullptr;
    if ((rc = g_ws.get(2, (size_t)n_fft * 16, &d_o)) != SFC_OK) return rc;
    MapParams m{};
    m.src = d_res;
    m.dst = d_o;
    m.total = n_fft;
    m.n = n_fft;
    m.inner = 1;
    m.mode = EM_IHFFT;
    cudaError_t e = launch_map(m, g_ws.stream);
    if (e != cudaSuccess) return cuda_fail(e, "ihfft reflection kernel");
    return download(out, d_o, (size_t)n_fft * 16);
}

// lib.rs:437-516: spectrum = fft(x, None) [padded to P = next_pow2(n)], first n bins * h, ifft(.., None)
// [n entries zero-padded to P again, scale 1/P, first n outputs].  h: 1 at DC (and Nyquist for even n),
// -2i on the positive frequencies, 0 on the negative ones.
SFC_EXPORT int sfc_hilbert(const double* x, int64_t n, double* out) {
    int rc;
    if ((rc = require_device()) != SFC_OK) return rc;
    if (!x || n <= 0 || !out) return fail(SFC_ERR_VALUE, "Input cannot be empty");
    const int64_t P = next_pow2_i64(n);
    void *d_x = nullptr, *d_h = nullptr, *d_s = nullptr, *d_o = nullptr;
    if ((rc = g_ws.get(0, (size_t)n * 8, &d_x)) != SFC_OK) return rc;
    if ((rc = g_ws.get(4, (size_t)n * 16, &d_h)) != SFC_OK) return rc;
    if ((rc = g_ws.get(1, (size_t)n * 16, &d_s)) != SFC_OK) return rc;
    if ((rc = g_ws.get(2, (size_t)n * 16, &d_o)) != SFC_OK) return rc;
    cudaStream_t st = g_ws.stream;
    std::vector<double> h(2 * (size_t)n, 0.0);
    const int64_t half = n % 2 == 0 ? n / 2 : (n + 1) / 2;
    h[0] = 1.0;
    for (int64_t k = 1; k < half; ++k) h[2 * k + 1] = -2.0;
    if (n % 2 == 0 && n / 2 < n) {
        h[2 * (n / 2)] = 1.0;  // Nyquist (for n == 2 this is index 1)
        h[2 * (n / 2) + 1] = 0.0;
    }
    cudaError_t e = cudaMemcpyAsync(d_x, x, (size_t)n * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_h, h.data(), (size_t)n * 16, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "H2D copy");
    e = cudaStreamSynchronize(st);  // h is a stack-lifetime host buffer
    if (e != cudaSuccess) return cuda_fail(e, "H2D copy");
    if ((rc = real_fft_crop(d_x, d_s, 1, n, 1, P, st)) != SFC_OK) return rc;
    MapParams m{};
    m.src = d_s;
    m.dst = d_s;
    m.tab = d_h;
    m.total = n;
    m.n = n;
    m.inner = 1;
    m.mode = EM_C_TAB;
    e = launch_map(m, st);
    if (e != cudaSuccess) return cuda_fail(e, "Hilbert filter kernel");
    sfc_desc d;
    memset(&d, 0, sizeof d);
    d.ndim = 1;
    d.shape[0] = P;
    d.naxes = 1;
    d.axes[0] = 0;
    d.kind = SFC_C2C;
    d.prec = SFC_PREC_F64;
    d.direction = SFC_INVERSE;
    d.scale = 1.0 / (double)P;
    d.flags = SFC_DESC_AXIS_LEN;
    d.axis_in_len = n;
    d.axis_out_len = n;
    PlanError perr{0, ""};
    std::shared_ptr<Plan> p = cached_plan(d, perr);
    if (!p) return fail(perr.code ? perr.code : SFC_ERR_PLAN, perr.msg);
    std::string es;
    rc = p->exec(d_s, d_o, st, es);
    if (rc != 0) return fail(rc, es);
    return download(out, d_o, (size_t)n * 16);
}

// spectrogram.rs:76-310 (stft) and :312-420 (spectrogram = a real-valued map of the one-sided stft).
// `window` holds the nperseg window samples (the caller evaluates get_window); boundary: 0 none, 1 "reflect",
// 2 "zeros", 3 "constant"; out_mode: sfc StftOut (0 complex [freq][frame], 1 psd, 2 magnitude, 3 phase, 4 angle).
SFC_EXPORT int sfc_stft(const double* x, int64_t len, const double* window, int64_t nperseg, int64_t noverlap, int64_t nfft,
                        int32_t detrend, int32_t onesided, int32_t boundary, int32_t out_mode, double scale, void* out,
                        int64_t out_cap_elems, int64_t* freq_len_out, int64_t* frames_out) {
    int rc;
    if ((rc = require_device()) != SFC_OK) return rc;
    if (!x || len <= 0) return fail(SFC_ERR_VALUE, "Input signal is empty");                    // :90-92
    if (nperseg <= 0 || !window) return fail(SFC_ERR_VALUE, "Segment length must be positive");  // :94-98
    if (nfft <= 0) nfft = nperseg;                                                              // :108
    if (nfft < nperseg) return fail(SFC_ERR_VALUE, "FFT length must be greater than or equal to segment length");
    if (noverlap < 0) noverlap = nperseg / 2;                                                   // :115
    if (noverlap >= nperseg) return fail(SFC_ERR_VALUE, "Overlap must be less than segment length");
    if (boundary < 0 || boundary > 3 || out_mode < 0 || out_mode > 4) return fail(SFC_ERR_VALUE, "unknown boundary / output mode");
    const int64_t step = nperseg - noverlap;
    const int64_t padded_len = boundary ? len + 2 * nperseg : len;
    if ((boundary && len < nperseg) || padded_len < nperseg)
        return fail(SFC_ERR_VALUE, "signal shorter than one segment (the reference underflows here)");
    const int64_t frames = 1 + (padded_len - nperseg) / step;  // :138, :160, :186
    const int64_t P = next_pow2_i64(nfft);                     // fft(&segment, None) pads to the next power of two (:290)
    const int64_t freq_len = onesided ? nfft / 2 + 1 : nfft;   // :193
    if (!onesided && P != nfft)
        return fail(SFC_ERR_VALUE, "two-sided stft needs a power-of-two nfft (the reference indexes out of bounds otherwise)");
    if (freq_len_out) *freq_len_out = freq_len;
    if (frames_out) *frames_out = frames;
    if (!out || out_cap_elems < freq_len * frames) return fail(SFC_ERR_VALUE, "output buffer too small");
    void *d_x = nullptr, *d_w = nullptr, *d_f = nullptr, *d_z = nullptr, *d_o = nullptr;
    const int64_t pitch = onesided ? P / 2 + 1 : P;
    const size_t out_es = out_mode == STFT_COMPLEX ? 16 : 8;
    if ((rc = g_ws.get(0, (size_t)len * 8, &d_x)) != SFC_OK) return rc;
    if ((rc = g_ws.get(4, (size_t)nperseg * 8, &d_w)) != SFC_OK) return rc;
    if ((rc = g_ws.get(3, (size_t)frames * P * 8, &d_f)) != SFC_OK) return rc;
    if ((rc = g_ws.get(1, (size_t)frames * pitch * 16, &d_z)) != SFC_OK) return rc;
    if ((rc = g_ws.get(2, (size_t)frames * freq_len * out_es, &d_o)) != SFC_OK) return rc;
    cudaStream_t st = g_ws.stream;
    cudaError_t e = cudaMemcpyAsync(d_x, x, (size_t)len * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_w, window, (size_t)nperseg * 8, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "H2D copy");
    FrameParams fp{};
    fp.x = (const double*)d_x;
    fp.win = (const double*)d_w;
    fp.dst = (double*)d_f;
    fp.len = len;
    fp.nperseg = nperseg;
    fp.step = step;
    fp.frames = frames;
    fp.P = P;
    fp.boundary = boundary;
    fp.detrend = detrend ? 1 : 0;
    e = launch_frames(fp, st);
    if (e != cudaSuccess) return cuda_fail(e, "stft framing kernel");
    sfc_desc d;
    memset(&d, 0, sizeof d);
    d.ndim = 2;
    d.shape[0] = frames;
    d.shape[1] = P;
    d.naxes = 1;
    d.axes[0] = 1;
    d.prec = SFC_PREC_F64;
    d.scale = 1.0;
    if (onesided) {
        d.kind = SFC_R2C;
    } else {
        d.kind = SFC_C2C;
        d.direction = SFC_FORWARD;
        d.flags = SFC_DESC_REAL_INPUT;
    }
    PlanError perr{0, ""};
    std::shared_ptr<Plan> p = cached_plan(d, perr);
    if (!p) return fail(perr.code ? perr.code : SFC_ERR_PLAN, perr.msg);
    std::string es;
    rc = p->exec(d_f, d_z, st, es);
    if (rc != 0) return fail(rc, es);
    StftOutParams op{};
    op.src = d_z;
    op.dst = d_o;
    op.frames = frames;
    op.src_pitch = pitch;
    op.freq_len = freq_len;
    op.mode = out_mode;
    op.scale = scale;
    e = launch_stft_out(op, st);
    if (e != cudaSuccess) return cuda_fail(e, "stft output kernel");
    return download(out, d_o, (size_t)frames * freq_len * out_es);
}

// ------------------------------------------------------------------------------------------------
// the segment loops of scirs2-signal (spectral.rs:346-395 welch, :580-616 stft, :186-207 periodogram).
// Rows x[f*step .. f*step + nperseg) -> detrend (0 none, 1 constant, 2 linear: spectral.rs:77-117) -> window -> zero-pad to P
// (the power of two `fft(&padded, None)` pads to) -> one batched real-to-complex plan -> the first `bins` bins.
// reduce = 0: out = complex f64 [frames][bins];  reduce = 1: out = f64 [bins] = scale * sum over frames of |X|^2.
SFC_EXPORT int sfc_signal_spectra(const double* x, int64_t len, const double* window, int64_t nperseg, int64_t step,
                                  int64_t frames, int64_t P, int32_t detrend, int32_t reduce, int64_t bins, double scale,
                                  void* out) {
    int rc;
    if ((rc = require_device()) != SFC_OK) return rc;
    if (!x || len <= 0) return fail(SFC_ERR_VALUE, "Input array is empty");
    if (nperseg <= 0 || !window || step <= 0 || frames <= 0) return fail(SFC_ERR_VALUE, "Segment length, step and count must be positive");
    if ((frames - 1) * step + nperseg > len) return fail(SFC_ERR_VALUE, "Not enough data points for given nperseg and noverlap");
    if (P < nperseg || !is_pow2_i64(P)) return fail(SFC_ERR_VALUE, "padded length must be a power of two >= nperseg");
    if (bins <= 0 || bins > P / 2 + 1) return fail(SFC_ERR_VALUE, "bins must be in 1 ..= P/2 + 1");
    if (detrend < 0 || detrend > 2 || reduce < 0 || reduce > 1 || !out) return fail(SFC_ERR_VALUE, "unknown detrend / reduce option");
    const int64_t pitch = P / 2 + 1;
    void *d_x = nullptr, *d_w = nullptr, *d_f = nullptr, *d_z = nullptr;
    if ((rc = g_ws.get(0, (size_t)len * 8, &d_x)) != SFC_OK) return rc;
    if ((rc = g_ws.get(4, (size_t)nperseg * 8, &d_w)) != SFC_OK) return rc;
    if ((rc = g_ws.get(3, (size_t)frames * P * 8, &d_f)) != SFC_OK) return rc;
    if ((rc = g_ws.get(1, (size_t)frames * pitch * 16, &d_z)) != SFC_OK) return rc;
    cudaStream_t st = g_ws.stream;
    cudaError_t e = cudaMemcpyAsync(d_x, x, (size_t)len * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_w, window, (size_t)nperseg * 8, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "H2D copy");
    FrameParams fp{};
    fp.x = (const double*)d_x;
    fp.win = (const double*)d_w;
    fp.dst = (double*)d_f;
    fp.len = len;
    fp.nperseg = nperseg;
    fp.step = step;
    fp.frames = frames;
    fp.P = P;
    fp.boundary = 0;
    fp.detrend = detrend;
    e = launch_frames(fp, st);
    if (e != cudaSuccess) return cuda_fail(e, "framing kernel");
    sfc_desc d;
    memset(&d, 0, sizeof d);
    d.ndim = 2;
    d.shape[0] = frames;
    d.shape[1] = P;
    d.naxes = 1;
    d.axes[0] = 1;
    d.prec = SFC_PREC_F64;
    d.scale = 1.0;
    d.kind = SFC_R2C;
    PlanError perr{0, ""};
    std::shared_ptr<Plan> p = cached_plan(d, perr);
    if (!p) return fail(perr.code ? perr.code : SFC_ERR_PLAN, perr.msg);
    std::string es;
    rc = p->exec(d_f, d_z, st, es);
    if (rc != 0) return fail(rc, es);
    if (reduce) {
        void *d_p = nullptr, *d_o = nullptr;
        const int64_t want = (frames + 63) / 64;
        const int32_t parts = (int32_t)(want < 1 ? 1 : (want > 1024 ? 1024 : want));
        if ((rc = g_ws.get(2, (size_t)parts * bins * 8, &d_p)) != SFC_OK) return rc;
        if ((rc = g_ws.get(5, (size_t)bins * 8, &d_o)) != SFC_OK) return rc;
        PsdSumParams sp{};
        sp.src = d_z;
        sp.partial = (double*)d_p;
        sp.dst = (double*)d_o;
        sp.frames = frames;
        sp.src_pitch = pitch;
        sp.bins = bins;
        sp.parts = parts;
        sp.scale = scale;
        e = launch_psd_sum(sp, st);
        if (e != cudaSuccess) return cuda_fail(e, "psd reduction kernel");
        return download(out, d_o, (size_t)bins * 8);
    }
    // rows of `pitch` complex entries on the device, rows of `bins` on the host
    e = cudaMemcpy2DAsync(out, (size_t)bins * 16, d_z, (size_t)pitch * 16, (size_t)bins * 16, (size_t)frames,
                          cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return cuda_fail(e, "D2H copy");
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(e, "transform execution");
    return SFC_OK;
}

// ------------------------------------------------------------------------------------------------
// memory_efficient.rs / ndim_optimized.rs — what benches/fft_benchmarks.rs:126-189 times.

// memory_efficient.rs:89-190.  n >= 32 takes the reference's "SIMD" branch (simd_support_available() is true on
// x86_64 / aarch64): fft_adaptive / ifft_adaptive, whose 1-D `norm` argument is ignored (simd_fft.rs:37-60) —
// forward is never scaled, inverse always by 1/n — and whose result has next_pow2(n) entries, so a non-power-of-two
// n indexes out of bounds there (ValueError here).  n < 32: rustfft directly, scale 1/n iff `normalize`.
SFC_EXPORT int sfc_fft_inplace(double* input, int64_t n, double* output, int64_t out_len, int32_t inverse, int32_t normalize) {
    int rc;
    if ((rc = require_device()) != SFC_OK) return rc;
    if (n <= 0 || !input) return fail(SFC_ERR_VALUE, "Input array is empty");
    if (!output || out_len < n) {
        char b[128];
        snprintf(b, sizeof b, "Output buffer is too small: got %lld, need %lld", (long long)out_len, (long long)n);
        return fail(SFC_ERR_VALUE, b);
    }
    if (n >= 32) {
        // fft(x, None) / ifft(x, None): padded to the next power of two; the forward result then has more than n
        // entries (out of bounds in the reference), the inverse one is truncated back to n (algorithms.rs:258-260)
        if (!inverse && !is_pow2_i64(n))
            return fail(SFC_ERR_VALUE, "fft_inplace (forward, n >= 32) needs a power-of-two length: the reference indexes out of bounds otherwise");
        int64_t got = 0;
        rc = inverse ? sfc_ifft(input, n, SFC_C128, -1, output, out_len, &got) : sfc_fft(input, n, SFC_C128, -1, output, out_len, &got);
        if (rc != SFC_OK) return rc;
    } else {
        const double scale = normalize ? 1.0 / (double)n : 1.0;
        void* d_res = nullptr;
        if ((rc = run_c2c_host(input, {n}, SFC_C128, {n}, {0}, inverse != 0, scale, &d_res)) != SFC_OK) return rc;
        if ((rc = download(output, d_res, (size_t)n * 16)) != SFC_OK) return rc;
    }
    memcpy(input, output, (size_t)n * 16);  // "copy the results back to the input and output buffers" (:121-124)
    return (int)std::min<int64_t>(n, 0x7fffffff);
}

// memory_efficient.rs:243-397: pad / crop to `shape`, rows then columns, 1/(rows*cols) iff normalize (either direction)
SFC_EXPORT int sfc_fft2_efficient(const void* x, int64_t rows, int64_t cols, int dtype, int64_t out_rows, int64_t out_cols,
                                  int32_t inverse, int32_t normalize, double* out) {
    int rc;
    if ((rc = require_device()) != SFC_OK) return rc;
    if (!dtype_ok(dtype)) return fail(SFC_ERR_VALUE, "unknown dtype");
    if (out_rows < 0) out_rows = rows;
    if (out_cols < 0) out_cols = cols;
    if (out_rows == 0 || out_cols == 0) return fail(SFC_ERR_VALUE, "Output dimensions must be positive");  // :256-260
    if (!x || rows <= 0 || cols <= 0 || !out) return fail(SFC_ERR_VALUE, "Input cannot be empty");
    const double scale = normalize ? 1.0 / ((double)out_rows * (double)out_cols) : 1.0;
    void* d_res = nullptr;
    if ((rc = run_c2c_host(x, {rows, cols}, dtype, {out_rows, out_cols}, {1, 0}, inverse != 0, scale, &d_res)) != SFC_OK) return rc;
    return download(out, d_res, (size_t)(out_rows * out_cols) * 16);
}

// memory_efficient.rs:401-580.  One transform of length n when it fits a chunk; otherwise — as the reference does —
// INDEPENDENT transforms of consecutive chunks (the last one shorter), concatenated; inverse chunks end up scaled by
// (1/len_chunk) * (chunk/n) (:566-577).  The full chunks run as one batched plan.
SFC_EXPORT int sfc_fft_streaming(const void* x, int64_t len, int dtype, int64_t n, int32_t inverse, int64_t chunk_size, double* out) {
    int rc;
    if ((rc = require_device()) != SFC_OK) return rc;
    if (!dtype_ok(dtype)) return fail(SFC_ERR_VALUE, "unknown dtype");
    if (!x || len <= 0 || !out) return fail(SFC_ERR_VALUE, "Input cannot be empty");
    const int64_t n_val = n > 0 ? n : len;
    int64_t chunk = chunk_size > 0 ? chunk_size : (len > 1000000 ? 1048576 : (len > 100000 ? 65536 : len));  // :412-424
    if (len <= chunk || n_val <= chunk) {  // :427
        void* d_res = nullptr;
        if ((rc = run_c2c_host(x, {std::min(len, n_val)}, dtype, {n_val}, {0}, inverse != 0, inverse ? 1.0 / (double)n_val : 1.0,
                               &d_res)) != SFC_OK)
            return rc;
        return download(out, d_res, (size_t)n_val * 16);
    }
    const int64_t full = n_val / chunk, rem = n_val - full * chunk;
    // full chunks: rows of a [full][chunk] array (input rows beyond `len` are zero: pad through the 2-D convert pass)
    const int64_t avail = std::min(len, n_val);
    void* d_res = nullptr;
    // upload what exists, widen to complex f64 [n_val] with zero fill (the 1-D convert / pad pass, no transform)
    if ((rc = run_c2c_host(x, {avail}, dtype, {n_val}, {}, false, 1.0, &d_res)) != SFC_OK) return rc;
    cudaStream_t st = g_ws.stream;
    void* d_out = nullptr;
    if ((rc = g_ws.get(2, (size_t)n_val * 16, &d_out)) != SFC_OK) return rc;
    auto run_rows = [&](int64_t rows_n, int64_t width, int64_t off, double scale) -> int {
        sfc_desc d;
        memset(&d, 0, sizeof d);
        d.ndim = 2;
        d.shape[0] = rows_n;
        d.shape[1] = width;
        d.naxes = 1;
        d.axes[0] = 1;
        d.kind = SFC_C2C;
        d.prec = SFC_PREC_F64;
        d.direction = inverse ? SFC_INVERSE : SFC_FORWARD;
        d.scale = scale;
        PlanError perr{0, ""};
        std::shared_ptr<Plan> p = cached_plan(d, perr);
        if (!p) return fail(perr.code ? perr.code : SFC_ERR_PLAN, perr.msg);
        std::string es2;
        const int r = p->exec((const char*)d_res + (size_t)off * 16, (char*)d_out + (size_t)off * 16, st, es2);
        if (r != 0) return fail(r, es2);
        return SFC_OK;
    };
    const double adj = (double)chunk / (double)n_val;  // full_scale / chunk_scale (:569-571)
    if ((rc = run_rows(full, chunk, 0, inverse ? (1.0 / (double)chunk) * adj : 1.0)) != SFC_OK) return rc;
    if (rem > 0 && (rc = run_rows(1, rem, full * chunk, inverse ? (1.0 / (double)rem) * adj : 1.0)) != SFC_OK) return rc;
    return download(out, d_out, (size_t)n_val * 16);
}

// ndim_optimized.rs:17-58: real input widened to complex; along every listed axis (sorted by stride, which commutes)
// `fft(&lane, None)` — padded to the next power of two — of which the first axis_len bins are written back (:73-82).
SFC_EXPORT int sfc_fftn_optimized(const double* x, int32_t ndim, const int64_t* shape, const int32_t* axes, int32_t naxes, double* out) {
    int rc;
    if ((rc = require_device()) != SFC_OK) return rc;
    if (!x || !out || !shape || ndim < 1 || ndim > SFC_MAX_DIMS) return fail(SFC_ERR_VALUE, "Input cannot be empty");
    std::vector<int64_t> sh(shape, shape + ndim);
    for (int64_t v : sh)
        if (v <= 0) return fail(SFC_ERR_VALUE, "Input cannot be empty");
    std::vector<int> ax;
    if (axes)
        ax.assign(axes, axes + naxes);
    else
        for (int i = 0; i < ndim; ++i) ax.push_back(i);
    for (int a : ax)
        if (a < 0 || a >= ndim) {
            char b[128];
            snprintf(b, sizeof b, "Axis %d is out of bounds for array with %d dimensions", a, ndim);
            return fail(SFC_ERR_VALUE, b);  // :135-142
        }
    std::stable_sort(ax.begin(), ax.end(), [](int p, int q) { return p > q; });  // smallest stride (last axis) first (:118-132)
    const int64_t total = vprod(sh);
    void *d_r = nullptr, *d_a = nullptr, *d_b = nullptr;
    if ((rc = g_ws.get(0, (size_t)total * 8, &d_r)) != SFC_OK) return rc;
    if ((rc = g_ws.get(1, (size_t)total * 16, &d_a)) != SFC_OK) return rc;
    if ((rc = g_ws.get(2, (size_t)total * 16, &d_b)) != SFC_OK) return rc;
    cudaStream_t st = g_ws.stream;
    cudaError_t e = cudaMemcpyAsync(d_r, x, (size_t)total * 8, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "H2D copy");
    const void* cur = d_r;
    bool cur_real = true;
    void* bufs[2] = {d_a, d_b};
    int which = 0;
    if (ax.empty()) {
        CopyParams c;
        memset(&c, 0, sizeof c);
        c.ndim = 1;
        c.dst_shape[0] = c.src_shape[0] = total;
        c.src_complex = 0;
        c.dst_complex = 1;
        c.src_f64 = c.dst_f64 = 1;
        c.scale = 1.0;
        c.total = total;
        c.src = d_r;
        c.dst = d_a;
        e = launch_nd_copy(c, st);
        if (e != cudaSuccess) return cuda_fail(e, "convert kernel");
        return download(out, d_a, (size_t)total * 16);
    }
    for (int a : ax) {
        int64_t O = 1, I = 1;
        for (int i = 0; i < a; ++i) O *= sh[i];
        for (int i = a + 1; i < ndim; ++i) I *= sh[i];
        sfc_desc d;
        memset(&d, 0, sizeof d);
        d.ndim = 3;
        d.shape[0] = O;
        d.shape[1] = next_pow2_i64(sh[a]);
        d.shape[2] = I;
        d.naxes = 1;
        d.axes[0] = 1;
        d.kind = SFC_C2C;
        d.prec = SFC_PREC_F64;
        d.direction = SFC_FORWARD;
        d.scale = 1.0;
        d.flags = SFC_DESC_AXIS_LEN | (cur_real ? SFC_DESC_REAL_INPUT : 0);
        d.axis_in_len = sh[a];
        d.axis_out_len = sh[a];
        PlanError perr{0, ""};
        std::shared_ptr<Plan> p = cached_plan(d, perr);
        if (!p) return fail(perr.code ? perr.code : SFC_ERR_PLAN, perr.msg);
        std::string es;
        rc = p->exec(cur, bufs[which], st, es);
        if (rc != 0) return fail(rc, es);
        cur = bufs[which];
        cur_real = false;
        which ^= 1;
    }
    return download(out, cur, (size_t)total * 16);
}

// ------------------------------------------------------------------------------------------------
// the chirp z-transform of czt.rs:45-275 with its FFT calls in place (the reference stubs them out with
// zero vectors, czt.rs:110-113, 239-252, so its czt returns zeros; this is the algorithm that file sets up):
//   X[k] = sum_j x[j] a^-j w^(jk) = wk2[k] * sum_j (x[j] a^-j wk2[j]) / wk2[k - j],   wk2[k] = w^(k^2/2)
// as a circular convolution of length nfft >= n + m - 1.  Two plan executions per call, everything else is tables:
// forward FFT with the input weights a^-j wk2[j] fused into the load and the chirp spectrum (pre-rotated so that the
// wanted window starts at 0) fused into the store; inverse FFT cropped to m outputs with wk2 fused into the store.
namespace {

struct CztTables {
    int64_t nfft = 0;
    void *d_awk2 = nullptr, *d_fwk2 = nullptr, *d_wk2c = nullptr;
};
std::mutex g_czt_mu;
std::map<std::tuple<int, int64_t, int64_t, int, double, double, double, double>, CztTables> g_czt;

int get_czt_tables(int64_t n, int64_t m, bool has_w, std::complex<double> w, std::complex<double> a, CztTables& out) {
    int dev = 0;
    cudaGetDevice(&dev);
    const auto key = std::make_tuple(dev, n, m, (int)has_w, w.real(), w.imag(), a.real(), a.imag());
    std::lock_guard<std::mutex> lk(g_czt_mu);
    auto it = g_czt.find(key);
    if (it != g_czt.end()) {
        out = it->second;
        return SFC_OK;
    }
    const int64_t mx = std::max(n, m);
    std::vector<std::complex<double>> wk2((size_t)mx);
    const double pi = 3.14159265358979323846;
    for (int64_t k = 0; k < mx; ++k) {
        if (has_w) {
            wk2[k] = std::pow(w, (double)k * (double)k / 2.0);  // czt.rs:85
        } else {  // czt.rs:89-95: exp(-i pi (k^2 mod 2m) / m)
            const double ph = -(pi * (double)((k * k) % (2 * m))) / (double)m;
            wk2[k] = std::polar(1.0, ph);
        }
    }
    CztTables t;
    t.nfft = next_pow2_i64(n + m - 1);  // the reference takes next_fast_len(n + m - 1): any length >= n + m - 1 gives the same result
    std::vector<std::complex<double>> awk2((size_t)n), chirp((size_t)t.nfft, 0.0), wk2c((size_t)m);
    for (int64_t k = 0; k < n; ++k) awk2[k] = std::pow(a, -(double)k) * wk2[k];  // czt.rs:103
    for (int64_t i = 1; i < n; ++i) chirp[n - 1 - i] = 1.0 / wk2[i];             // czt.rs:108-113
    for (int64_t i = 0; i < m; ++i) chirp[n - 1 + i] = 1.0 / wk2[i];
    for (int64_t k = 0; k < m; ++k) wk2c[k] = std::conj(wk2[k]);  // the inverse plan multiplies by conj(table) (conjugation trick)
    // spectrum of the reciprocal chirp with our own transform
    void* d_res = nullptr;
    int rc = run_c2c_host(chirp.data(), {t.nfft}, SFC_C128, {t.nfft}, {0}, false, 1.0, &d_res);
    if (rc != SFC_OK) return rc;
    std::vector<std::complex<double>> F((size_t)t.nfft);
    if ((rc = download(F.data(), d_res, (size_t)t.nfft * 16)) != SFC_OK) return rc;
    // rotate so that y[n-1 .. n-1+m) of the convolution lands at [0, m): multiply bin k by exp(+2 pi i (n-1) k / nfft)
    for (int64_t k = 0; k < t.nfft; ++k) {
        const int64_t r = ((n - 1) * k) % t.nfft;
        F[k] *= std::polar(1.0, 2.0 * pi * (double)r / (double)t.nfft);
    }
    cudaError_t e = cudaMalloc(&t.d_awk2, (size_t)n * 16);
    if (e == cudaSuccess) e = cudaMalloc(&t.d_fwk2, (size_t)t.nfft * 16);
    if (e == cudaSuccess) e = cudaMalloc(&t.d_wk2c, (size_t)m * 16);
    if (e == cudaSuccess) e = cudaMemcpy(t.d_awk2, awk2.data(), (size_t)n * 16, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(t.d_fwk2, F.data(), (size_t)t.nfft * 16, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(t.d_wk2c, wk2c.data(), (size_t)m * 16, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return cuda_fail(e, "CZT tables");
    g_czt[key] = t;
    out = t;
    return SFC_OK;
}

}  // namespace

// x: [rows][n] complex f64 -> out: [rows][m] complex f64; has_w == 0: w = exp(-2 pi i / m) (czt.rs:87-96)
SFC_EXPORT int sfc_czt(const double* x, int64_t rows, int64_t n, int64_t m, int32_t has_w, double w_re, double w_im, double a_re,
                       double a_im, double* out) {
    int rc;
    if ((rc = require_device()) != SFC_OK) return rc;
    if (n < 1) return fail(SFC_ERR_VALUE, "n must be positive");  // czt.rs:70-72
    if (m < 1) return fail(SFC_ERR_VALUE, "m must be positive");  // czt.rs:75-77
    if (!x || !out || rows < 1) return fail(SFC_ERR_VALUE, "Input cannot be empty");
    CztTables t;
    if ((rc = get_czt_tables(n, m, has_w != 0, {w_re, w_im}, {a_re, a_im}, t)) != SFC_OK) return rc;
    void *d_x = nullptr, *d_s = nullptr, *d_o = nullptr;
    if ((rc = g_ws.get(0, (size_t)rows * n * 16, &d_x)) != SFC_OK) return rc;
    if ((rc = g_ws.get(1, (size_t)rows * t.nfft * 16, &d_s)) != SFC_OK) return rc;
    if ((rc = g_ws.get(2, (size_t)rows * m * 16, &d_o)) != SFC_OK) return rc;
    cudaStream_t st = g_ws.stream;
    cudaError_t e = cudaMemcpyAsync(d_x, x, (size_t)rows * n * 16, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "H2D copy");
    sfc_desc d;
    memset(&d, 0, sizeof d);
    d.ndim = 2;
    d.shape[0] = rows;
    d.shape[1] = t.nfft;
    d.naxes = 1;
    d.axes[0] = 1;
    d.kind = SFC_C2C;
    d.prec = SFC_PREC_F64;
    d.direction = SFC_FORWARD;
    d.scale = 1.0;
    d.flags = SFC_DESC_AXIS_LEN | SFC_DESC_AUX_MUL;
    d.axis_in_len = n;
    d.axis_out_len = t.nfft;
    d.aux_in = t.d_awk2;
    d.aux_out = t.d_fwk2;
    PlanError perr{0, ""};
    std::shared_ptr<Plan> pf = cached_plan(d, perr);
    if (!pf) return fail(perr.code ? perr.code : SFC_ERR_PLAN, perr.msg);
    d.direction = SFC_INVERSE;
    d.scale = 1.0 / (double)t.nfft;
    d.axis_in_len = t.nfft;
    d.axis_out_len = m;
    d.aux_in = nullptr;
    d.aux_out = t.d_wk2c;
    std::shared_ptr<Plan> pi = cached_plan(d, perr);
    if (!pi) return fail(perr.code ? perr.code : SFC_ERR_PLAN, perr.msg);
    std::string es;
    if ((rc = pf->exec(d_x, d_s, st, es)) != 0) return fail(rc, es);
    if ((rc = pi->exec(d_s, d_o, st, es)) != 0) return fail(rc, es);
    return download(out, d_o, (size_t)rows * m * 16);
}
