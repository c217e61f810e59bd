// plan.cu — planner + executor.  See plan.h.
#include "plan.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdlib>
#include <cstring>
#include <map>
#include <tuple>
#ifndef SFC_HOST_EMUL
#include <cuda.h>  // CUtensorMap + enums only: the encoder is fetched with cudaGetDriverEntryPoint, libcuda is not linked
#endif

namespace sfc {

// ------------------------------------------------------------ kernel table

// ---- planner options: environment variables that can be overridden at run time (sfc_planner_set_option), so that a tuner
// can time variants of a plan inside one process and persist the winners (auto_tuning.rs:188-229)
static std::mutex g_knob_mu;
static std::map<std::string, std::string>& knob_overrides() {
    static std::map<std::string, std::string> m;
    return m;
}
static thread_local std::string g_knob_tmp;
static const char* knob_env(const char* name) {
    {
        std::lock_guard<std::mutex> lk(g_knob_mu);
        auto it = knob_overrides().find(name);
        if (it != knob_overrides().end()) {
            g_knob_tmp = it->second;
            return g_knob_tmp.c_str();
        }
    }
    return getenv(name);
}
void planner_set_option(const char* name, const char* value) {
    std::lock_guard<std::mutex> lk(g_knob_mu);
    if (value) knob_overrides()[name] = value;
    else knob_overrides().erase(name);
}
std::string planner_get_option(const char* name) {
    const char* v = knob_env(name);
    return v ? std::string(v) : std::string();
}

static std::vector<KernelEntry>& ktable() {
    static std::vector<KernelEntry> v;
    return v;
}
static void add_entry(const KernelEntry& e) { ktable().push_back(e); }
static std::once_flag g_kernels_once;

const KernelEntry* kernel_table(int* count) {
    std::call_once(g_kernels_once, [] {
        register_kernels_f64_small(add_entry);
        register_kernels_f64_mid(add_entry);
        register_kernels_f64_big(add_entry);
        register_kernels_f64_dbl_a(add_entry);
        register_kernels_f64_dbl_b(add_entry);
        register_kernels_f32_small(add_entry);
        register_kernels_f32_mid(add_entry);
        register_kernels_f32_big(add_entry);
        register_kernels_f32_dbl_a(add_entry);
        register_kernels_f32_dbl_b(add_entry);
        register_kernels_e8(add_entry);
        register_kernels_f64_real(add_entry);
        register_kernels_f32_real(add_entry);
        register_kernels_pipe(add_entry);
        register_kernels_dct(add_entry);
        register_kernels_pipe_dbl(add_entry);
        register_kernels_r3(add_entry);
    });
    if (count) *count = (int)ktable().size();
    return ktable().data();
}

const KernelEntry* find_kernel(int prec, int L, int TL, int dbl, int mode) {
    int n = 0;
    const KernelEntry* t = kernel_table(&n);
    for (int i = 0; i < n; ++i)
        if (t[i].prec == prec && t[i].L == L && t[i].TL == TL && t[i].dbl == dbl && t[i].mode == mode && t[i].groups == 1)
            return &t[i];
    return nullptr;
}

// same tile shape, different flavour (nullptr when that flavour is not compiled)
static const KernelEntry* flavour_of(const KernelEntry* k, int mode) {
    int n = 0;
    const KernelEntry* t = kernel_table(&n);
    for (int i = 0; i < n; ++i)
        if (t[i].prec == k->prec && t[i].L == k->L && t[i].TL == k->TL && t[i].dbl == k->dbl && t[i].E == k->E &&
            t[i].groups == k->groups && t[i].mode == mode)
            return &t[i];
    return nullptr;
}

static bool fast_enabled() {
    const int v = [] {
        const char* e = knob_env("SFC_FAST");
        return e ? atoi(e) : 1;
    }();
    return v != 0;
}

// ROW tiles want few lanes per CTA (small tiles, more CTAs per SM); COL tiles want
// many adjacent lanes (>= 128 B contiguous per element row).
static int groups_mode() {
    const int v = [] {
        const char* e = knob_env("SFC_GROUPS");
        return e ? atoi(e) : 0;  // measured on B200: two named-barrier groups per CTA are slower (1024^3: 69% -> 61%)
    }();
    return v;
}

// one named-barrier group per lane for the contiguous-row tiles whose lanes are whole warps
static int row_lane_groups() {
#ifdef SFC_HOST_EMUL
    return 0;  // tests/emul has no named barriers: the host emulation keeps the single-group flavour (same arithmetic)
#endif
    const int v = [] {
        const char* e = knob_env("SFC_ROW_LANE_GROUPS");
        return e ? atoi(e) : 1;  // measured (sustained): 512-point rows 93.2 -> 97.8 %, 1024-point rows 92.4 -> 93.6 %
    }();
    return v;
}

static int col_tl_cap() {
    const int v = [] {
        const char* e = knob_env("SFC_COL_TL");
        return e ? atoi(e) : 0;
    }();
    return v;
}

// column tiles of internal passes: widest tile whose exchange buffer stays under this many KiB (100 = two 64 KiB
// tiles per SM, 40 = four 32 KiB tiles per SM)
static int col_smem_cap_kb() {
    const int v = [] {
        const char* e = knob_env("SFC_COL_SMEM_KB");
        return e ? atoi(e) : 100;
    }();
    return v;
}
static int forced_e() {
    const int v = [] {
        const char* e = knob_env("SFC_FORCE_E");
        return e ? atoi(e) : 0;
    }();
    return v;
}

static const KernelEntry* pick_kernel(int prec, int L, bool want_wide, int dbl) {
    int n = 0;
    const KernelEntry* t = kernel_table(&n);
    const KernelEntry* best = nullptr;
    const int want_e = forced_e() ? forced_e() : 16;
    bool have_e = false;
    for (int i = 0; i < n; ++i)
        if (t[i].mode == 0 && t[i].prec == prec && t[i].L == L && t[i].dbl == dbl && t[i].E == std::min(want_e, L))
            have_e = true;
    const int cap = (want_wide && col_tl_cap() > 0) ? col_tl_cap() : (1 << 30);
    const KernelEntry* smallest = nullptr;
    for (int i = 0; i < n; ++i) {
        if (t[i].mode != 0 || t[i].groups != 1 || t[i].prec != prec || t[i].L != L || t[i].dbl != dbl) continue;
        if (have_e ? t[i].E != std::min(want_e, L) : t[i].E != std::min(16, L)) continue;
        if (!smallest || t[i].TL < smallest->TL) smallest = &t[i];
        if (t[i].TL > cap) continue;
        if (!best || (want_wide ? t[i].TL > best->TL : t[i].TL < best->TL)) best = &t[i];
    }
    if (!best) best = smallest;
    if (best && want_wide && groups_mode() >= 1) {
        // same tile, two independent thread groups (overlaps one group's loads with the other's math)
        for (int i = 0; i < n; ++i)
            if (t[i].mode == 0 && t[i].groups == 2 && t[i].prec == prec && t[i].L == L && t[i].TL == best->TL &&
                t[i].dbl == dbl && t[i].E == best->E)
                return &t[i];
    }
    return best;
}

// Internal column passes of four-step / Bluestein: the widest tile whose exchange buffer still
// lets two CTAs share an SM (measured: 2^20 four-step 60% -> 77% of HBM peak vs one 128 KiB CTA).
static const KernelEntry* pick_kernel_two_per_sm(int prec, int L, int dbl) {
    int n = 0;
    const KernelEntry* t = kernel_table(&n);
    // the forced points-per-thread variant only where it is compiled, else the default (16)
    int want_e = std::min(16, L);
    if (forced_e())
        for (int i = 0; i < n; ++i)
            if (t[i].mode == 0 && t[i].groups == 1 && t[i].prec == prec && t[i].L == L && t[i].dbl == dbl &&
                t[i].E == std::min(forced_e(), L))
                want_e = t[i].E;
    const KernelEntry *best = nullptr, *smallest = nullptr;
    for (int i = 0; i < n; ++i) {
        if (t[i].mode != 0 || t[i].groups != 1 || t[i].prec != prec || t[i].L != L || t[i].dbl != dbl || t[i].E != want_e)
            continue;
        if (!smallest || t[i].TL < smallest->TL) smallest = &t[i];
        if (t[i].smem > (size_t)col_smem_cap_kb() * 1024) continue;
        if (!best || t[i].TL > best->TL) best = &t[i];
    }
    if (groups_mode() >= 2) {
        // grouped wide tile instead of the narrow two-per-SM tile
        for (int i = 0; i < n; ++i)
            if (t[i].mode == 0 && t[i].groups == 2 && t[i].prec == prec && t[i].L == L && t[i].dbl == dbl) return &t[i];
    }
    return best ? best : smallest;
}

// ------------------------------------------------------------- device tables

using TableKey = std::tuple<int, int, int, int64_t, int64_t, int64_t, int64_t>;
static std::recursive_mutex g_table_mu;
static std::map<TableKey, const void*>& tables() {
    static std::map<TableKey, const void*> m;
    return m;
}
enum TableKind { TK_STAGE = 1, TK_RTW, TK_FS_LO, TK_FS_HI, TK_CHIRP, TK_BLUE, TK_CR_LO, TK_CR_HI, TK_DCT2, TK_DCT4_PRE, TK_DCT4_POST };

static inline void unit_root(long double num, long double den, long double& c, long double& s) {
    // exp(-2*pi*i*num/den) in extended precision
    const long double two_pi = 6.283185307179586476925286766559005768L;
    const long double a = two_pi * (num / den);
    c = cosl(a);
    s = -sinl(a);
}

static const void* upload_table(const std::vector<long double>& re, const std::vector<long double>& im,
                                int prec, PlanError& err) {
    const size_t n = re.size();
    void* d = nullptr;
    const size_t bytes = n * (prec == PREC_F64 ? 16 : 8);
    if (cudaMalloc(&d, bytes ? bytes : 16) != cudaSuccess) {
        err = {SFC_ERR_MEMORY, "cudaMalloc failed for a twiddle table"};
        cudaGetLastError();
        return nullptr;
    }
    cudaError_t e;
    if (prec == PREC_F64) {
        std::vector<double> h(2 * n);
        for (size_t i = 0; i < n; ++i) {
            h[2 * i] = (double)re[i];
            h[2 * i + 1] = (double)im[i];
        }
        e = cudaMemcpy(d, h.data(), bytes, cudaMemcpyHostToDevice);
    } else {
        std::vector<float> h(2 * n);
        for (size_t i = 0; i < n; ++i) {
            h[2 * i] = (float)re[i];
            h[2 * i + 1] = (float)im[i];
        }
        e = cudaMemcpy(d, h.data(), bytes, cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
        err = {SFC_ERR_BACKEND, std::string("table upload failed: ") + cudaGetErrorString(e)};
        cudaFree(d);
        return nullptr;
    }
    return d;
}

static int cur_device() {
    int d = 0;
    cudaGetDevice(&d);
    return d;
}

static const void* roots_table(int kind, int prec, int64_t count, long double den, long double mult,
                               int64_t keyN, PlanError& err) {
    std::lock_guard<std::recursive_mutex> lk(g_table_mu);
    TableKey key{cur_device(), kind, prec, keyN, count, (int64_t)mult, 0};
    auto it = tables().find(key);
    if (it != tables().end()) return it->second;
    std::vector<long double> re(count), im(count);
    for (int64_t j = 0; j < count; ++j) unit_root((long double)j * mult, den, re[j], im[j]);
    const void* d = upload_table(re, im, prec, err);
    if (d) tables()[key] = d;
    return d;
}

// DCT-IV load twiddles exp(-i pi (4j+1) / (4N)), j < N/2
static const void* table_dct4_pre(int prec, int64_t N, PlanError& err) {
    std::lock_guard<std::recursive_mutex> lk(g_table_mu);
    TableKey key{cur_device(), TK_DCT4_PRE, prec, N, N / 2, 0, 0};
    auto it = tables().find(key);
    if (it != tables().end()) return it->second;
    std::vector<long double> re(N / 2), im(N / 2);
    for (int64_t j = 0; j < N / 2; ++j) unit_root((long double)(4 * j + 1), 8.0L * (long double)N, re[j], im[j]);
    const void* d = upload_table(re, im, prec, err);
    if (d) tables()[key] = d;
    return d;
}

const void* table_stage_tw(int prec, int L, PlanError& err) {
    return roots_table(TK_STAGE, prec, L, (long double)L, 1.0L, L, err);
}

const void* table_rtw(int prec, int L, PlanError& err) {
    const int cnt = std::max(L / 8, 1);  // covers E = 16 (L/16 used) and E = 8
    return roots_table(TK_RTW, prec, cnt, 2.0L * L, 1.0L, L, err);
}

bool table_fourstep(int prec, int64_t M, const void** lo, const void** hi, int* shift, PlanError& err) {
    int lg = 0;
    while (((int64_t)1 << lg) < M) ++lg;
    const int sh = (lg + 1) / 2;
    const int64_t nlo = (int64_t)1 << sh;
    const int64_t nhi = std::max<int64_t>((M + nlo - 1) >> sh, 1);  // covers (M-1) >> sh for any M (equal to M >> sh for powers of two)
    *lo = roots_table(TK_FS_LO, prec, nlo, (long double)M, 1.0L, M, err);
    if (!*lo) return false;
    *hi = roots_table(TK_FS_HI, prec, nhi, (long double)M, (long double)nlo, M, err);
    if (!*hi) return false;
    *shift = sh;
    return true;
}

// two-level table of R(k) = exp(-i*pi*k/N), k < 2N (always f64): R(k) = hi[k >> shift] * lo[k & mask]
bool table_chirp_roots(int64_t N, const void** lo, const void** hi, int* shift, PlanError& err) {
    int lg = 0;
    while (((int64_t)1 << lg) < 2 * N) ++lg;
    const int sh = (lg + 1) / 2;
    const int64_t nlo = (int64_t)1 << sh;
    const int64_t nhi = (2 * N + nlo - 1) / nlo;
    *lo = roots_table(TK_CR_LO, PREC_F64, nlo, (long double)(2 * N), 1.0L, N, err);
    if (!*lo) return false;
    *hi = roots_table(TK_CR_HI, PREC_F64, nhi, (long double)(2 * N), (long double)nlo, N, err);
    if (!*hi) return false;
    *shift = sh;
    return true;
}

static bool chirp_gen_enabled() {
    const int v = [] {
        const char* e = knob_env("SFC_CHIRP_GEN");
        return e ? atoi(e) : 1;
    }();
    return v != 0;
}

const void* table_chirp(int prec, int64_t N, PlanError& err) {
    std::lock_guard<std::recursive_mutex> lk(g_table_mu);
    TableKey key{cur_device(), TK_CHIRP, prec, N, 0, 0, 0};
    auto it = tables().find(key);
    if (it != tables().end()) return it->second;
    std::vector<long double> re(N), im(N);
    const unsigned __int128 twoN = 2 * (unsigned __int128)N;
    for (int64_t n = 0; n < N; ++n) {
        // exp(-i*pi*n^2/N) with the phase reduced exactly: n^2 mod 2N
        const unsigned __int128 r = ((unsigned __int128)n * (unsigned __int128)n) % twoN;
        unit_root((long double)(uint64_t)r, (long double)(2 * N), re[n], im[n]);
    }
    const void* d = upload_table(re, im, prec, err);
    if (d) tables()[key] = d;
    return d;
}

const void* table_bluestein_b(int prec, int64_t N, int64_t M, int64_t L1, int64_t L2, PlanError& err, int64_t L3) {
    std::lock_guard<std::recursive_mutex> lk(g_table_mu);
    TableKey key{cur_device(), TK_BLUE, prec, N, M, L1, L2 + (L3 << 32)};
    auto it = tables().find(key);
    if (it != tables().end()) return it->second;

    // b[n] = conj(chirp[n]) for |n| < N (wrapped into length M), FFT_M(b)/M computed with
    // our own f64 transform, then laid out for the consuming pass.
    std::vector<double> hb(2 * (size_t)M, 0.0);
    const unsigned __int128 twoN = 2 * (unsigned __int128)N;
    for (int64_t n = 0; n < N; ++n) {
        const unsigned __int128 r = ((unsigned __int128)n * (unsigned __int128)n) % twoN;
        long double c, s;
        unit_root((long double)(uint64_t)r, (long double)(2 * N), c, s);
        hb[2 * n] = (double)c;
        hb[2 * n + 1] = (double)(-s);
        if (n > 0) {
            hb[2 * (M - n)] = (double)c;
            hb[2 * (M - n) + 1] = (double)(-s);
        }
    }
    sfc_desc d{};
    d.ndim = 1;
    d.shape[0] = M;
    d.naxes = 1;
    d.axes[0] = 0;
    d.kind = SFC_C2C;
    d.prec = SFC_PREC_F64;
    d.direction = SFC_FORWARD;
    d.scale = 1.0 / (double)M;
    std::shared_ptr<Plan> pl = Plan::create(d, err);
    if (!pl) return nullptr;
    void* dbuf = nullptr;
    const size_t bytes = (size_t)M * 16;
    if (cudaMalloc(&dbuf, bytes) != cudaSuccess) {
        cudaGetLastError();
        err = {SFC_ERR_MEMORY, "cudaMalloc failed for the Bluestein kernel spectrum"};
        return nullptr;
    }
    std::string es;
    cudaError_t e = cudaMemcpy(dbuf, hb.data(), bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && pl->exec(dbuf, dbuf, 0, es) != 0) {
        err = {SFC_ERR_COMPUTATION, "Bluestein kernel spectrum: " + es};
        cudaFree(dbuf);
        return nullptr;
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    if (e == cudaSuccess) e = cudaMemcpy(hb.data(), dbuf, bytes, cudaMemcpyDeviceToHost);
    cudaFree(dbuf);
    if (e != cudaSuccess) {
        err = {SFC_ERR_BACKEND, std::string("Bluestein kernel spectrum: ") + cudaGetErrorString(e)};
        return nullptr;
    }
    std::vector<long double> re(M), im(M);
    if (L3 > 0) {
        // three-level work order: position k1*(L2*L3) + k2*L3 + k3 holds bin k = k1 + L1*(k2 + L2*k3)
        for (int64_t k1 = 0; k1 < L1; ++k1)
            for (int64_t k2 = 0; k2 < L2; ++k2)
                for (int64_t k3 = 0; k3 < L3; ++k3) {
                    const int64_t k = k1 + L1 * (k2 + L2 * k3);
                    const int64_t pos = k1 * (L2 * L3) + k2 * L3 + k3;
                    re[pos] = hb[2 * k];
                    im[pos] = hb[2 * k + 1];
                }
    } else if (L1 > 0) {
        for (int64_t k1 = 0; k1 < L1; ++k1)
            for (int64_t k2 = 0; k2 < L2; ++k2) {
                const int64_t k = k1 + L1 * k2;
                re[k1 * L2 + k2] = hb[2 * k];
                im[k1 * L2 + k2] = hb[2 * k + 1];
            }
    } else {
        for (int64_t k = 0; k < M; ++k) {
            re[k] = hb[2 * k];
            im[k] = hb[2 * k + 1];
        }
    }
    const void* t = upload_table(re, im, prec, err);
    if (t) tables()[key] = t;
    return t;
}

// ------------------------------------------------------------------ builder

static inline bool is_pow2(int64_t n) { return n > 0 && (n & (n - 1)) == 0; }
static inline int ilog3_exact(int64_t n) {  // k if n == 3^k, else -1
    int k = 0;
    while (n > 1 && n % 3 == 0) {
        n /= 3;
        ++k;
    }
    return n == 1 ? k : -1;
}
// power-of-three tiles (r3_tile.cuh): 1 = on (default), 0 = every non-power-of-two length goes through Bluestein
static bool radix3_enabled() {
    const int v = [] {
        const char* e = knob_env("SFC_RADIX3");
        return e ? atoi(e) : 1;
    }();
    return v != 0;
}
// rows: the 243-thread tile (three CTAs per SM); strided passes: the 486-thread tile with twice the lanes (longer segments).
// Measured (profiles/r2g_r3_tiles_latency_tmap.log): contiguous 729-point rows 61.9 % (729 x 9 lanes) -> 77.0 % (729 x 3),
// 2187-point rows 54.1 % -> 67.6 %; in the strided passes of 3^13 narrower tiles lose.
static const KernelEntry* pick_r3(int prec, int L, bool strided) {
    int n = 0;
    const KernelEntry* t = kernel_table(&n);
    {   // experiment knob: SFC_R3_TL_<L>=<lanes per tile>
        char nm[32];
        snprintf(nm, sizeof nm, "SFC_R3_TL_%d", L);
        const char* e = knob_env(nm);
        if (e)
            for (int i = 0; i < n; ++i)
                if (t[i].mode == 10 && t[i].prec == prec && t[i].L == L && t[i].TL == atoi(e)) return &t[i];
    }
    const KernelEntry* best = nullptr;
    const int want = strided ? 486 : 243;
    for (int i = 0; i < n; ++i) {
        if (t[i].mode != 10 || t[i].prec != prec || t[i].L != L) continue;
        if (!best || std::abs(t[i].threads - want) < std::abs(best->threads - want)) best = &t[i];
    }
    return best;
}
static inline int64_t next_pow2(int64_t n) {
    int64_t p = 1;
    while (p < n) p <<= 1;
    return p;
}
static inline int ilog2_64(int64_t n) {
    int l = 0;
    while (((int64_t)1 << l) < n) ++l;
    return l;
}

static int col_single_limit(int prec) {
    const int v = [] {
        const char* e = knob_env("SFC_COL_SINGLE_MAX");
        return e ? atoi(e) : 0;
    }();
    if (v > 0) return v;
    return prec == PREC_F64 ? 2048 : 4096;
}

// row length that is transformed as two half-length lanes (0 = never); SFC_ROW_SPLIT overrides
static int row_split_len(int prec) {
    const int v = [] {
        const char* e = knob_env("SFC_ROW_SPLIT");
        return e ? atoi(e) : -1;
    }();
    if (v >= 0) return v;
    (void)prec;
    return 0;  // measured on B200: 8192-point f64 rows 52.3 % split vs 52.8 % single tile (interleaved 16 B stores)
}

// L2 blocking of multi-pass transforms (four-step, Bluestein): a round covers at most this many
// bytes of work area so that pass k+1 reads what pass k wrote out of the 126 MB L2, and `ways`
// rounds run concurrently on side streams to keep every SM busy.  0 = off.
static int64_t l2_chunk_bytes() {
    const int64_t v = [] {
        const char* e = knob_env("SFC_L2_CHUNK_MB");
        // measured on B200 (2^20 x 64 four-step, Bluestein 1,000,003 x 32): no gain from L2-resident
        // rounds (75.6 % either way / 56 % vs 61 %) — the 64 KiB-tile passes are SM-bound, not DRAM-bound
        int64_t mb = e ? atoll(e) : 0;
        return mb << 20;
    }();
    return v;
}
static int l2_ways() {
    const int v = [] {
        const char* e = knob_env("SFC_L2_WAYS");
        int w = e ? atoi(e) : 3;
        return std::max(1, std::min(w, 16));
    }();
    return v;
}
static int64_t l2_max_batch_bytes() {
    const int64_t v = [] {
        const char* e = knob_env("SFC_L2_MAXB_MB");
        return e ? (atoll(e) << 20) : l2_chunk_bytes();
    }();
    return v;
}
static int64_t l2_total_bytes() {
    const int64_t v = [] {
        const char* e = knob_env("SFC_L2_TOTAL_MB");
        int64_t mb = e ? atoll(e) : 72;
        return mb << 20;
    }();
    return v;
}

// measured (rows, 4 GiB, f64 / f32): L = 16: 61 -> 103 % / 30 -> 96 % of the measured HBM peak, L = 8: 58 -> 80 % /
// 53 -> 73 %; L = 2, 4 (32..64 B rows, already coalesced) lose 3-14 points; L = 32 / 64 (which also exchange through
// the buffer) 69 -> 69 % / 89 -> 73 %: staged for 8 <= L <= 16 only
static int stage_io_max_len() {
    const int v = [] {
        const char* e = knob_env("SFC_STAGE_IO");
        return e ? atoi(e) : 16;
    }();
    return v;
}
static int pipe_enabled() {  // 0 off, 1 row and column tiles, 2 row tiles only
    const int v = [] {
        // measured on B200 (profiles/README.md): the half-size split exchange that makes room for the landing
        // buffer costs more (FFT phase 4.7k -> 6.6k cycles per 4096-point tile) than the hidden load latency
        // gains (c2c 92 % -> 88 %, Bluestein 59 % -> 57 %); per-row bulk copies for column tiles are far too
        // slow (60 cycles each).  Kept selectable for the next round's tensor-map variant.
        const char* e = knob_env("SFC_PIPE");
        return e ? atoi(e) : 0;
    }();
    return v;
}
// measured on B200: parity identical, c2c 65,536 x 4096 93 -> 89 %, 1024 / 2048-point rows 98 -> 88 % / 88 -> 82 %: with ONE
// landing buffer per SM (shared memory has no room for two next to two full-size exchange buffers) only 64 KiB of loads are
// in flight per SM, less than the two independent CTAs of the plain flavour keep in flight.  Off.
static bool gpipe_enabled() {
    const int v = [] {
        const char* e = knob_env("SFC_GPIPE");
        return e ? atoi(e) : 0;
    }();
    return v != 0;
}
static int64_t gpipe_min_tiles() {
    const int64_t v = [] {
        const char* e = knob_env("SFC_GPIPE_MIN_TILES");
        return e ? atoll(e) : 8 * 148;  // four tiles for each group of each of the 148 persistent CTAs
    }();
    return v;
}
// late-prefetch persistent flavour (TM_PIPE_LATE).  Measured on B200 (profiles/r2c_sustained_ab.log, r2d_copy_power_tmap.log):
//   contiguous rows (1-D bulk copies): c2c 65,536 x 4096 79.3 % -> 75.3 % sustained, 8192-point rows 58.6 % -> 55.6 %,
//   2048-point rows 85.2 % -> 82.9 %: the extra shared-memory read of the landed tile costs more than the hidden latency;
//   strided tiles through a tensor map (UTMALDG): 512 x 8 tiles (128 B segments) lose (fftn 512^3 91.7 % -> 83.5 %), the
//   1024 x 4 tiles of the four-step path (64 B segments, where per-thread loads are least efficient) gain (2^20 x 64:
//   75.0 % -> 78.1 %).
// 0 off | 1 row tiles <= 64 KiB | 2 + 128 KiB row tiles | 3 + every strided tile | 4 (default) only the narrow strided tiles
// (<= 64 B segments) and multi-row tiles of multi-pass (four-step) plans.
static int pipe_late_enabled() {
    const int v = [] {
        const char* e = knob_env("SFC_PIPE_LATE");
        return e ? atoi(e) : 4;
    }();
    return v;
}
static int64_t pipe_late_min_tiles() {
    const int64_t v = [] {
        const char* e = knob_env("SFC_PIPE_LATE_MIN_TILES");
        return e ? atoll(e) : 4 * 148;  // two tiles for each of the 296 resident CTAs, or prefetching never pays
    }();
    return v;
}
static bool pipe_big_enabled() {
    const int v = [] {
        const char* e = knob_env("SFC_PIPE_BIG");
        return e ? atoi(e) : 1;
    }();
    return v != 0;
}
static int64_t pipe_min_tiles() {
    const int64_t v = [] {
        const char* e = knob_env("SFC_PIPE_MIN_TILES");
        return e ? atoll(e) : 1184;  // 4 tiles for each of the 296 resident CTAs
    }();
    return v;
}
static int tile_group_log2() {
    const int v = [] {
        const char* e = knob_env("SFC_TILE_GROUP_LOG2");
        return e ? std::max(0, std::min(atoi(e), 10)) : 3;  // table-driven Bluestein 57 % -> 61 %; neutral with generated chirps
    }();
    return v;
}
static int blue_l1_cap() {
    const int v = [] {
        const char* e = knob_env("SFC_BLUE_L1");
        return e ? atoi(e) : 1024;
    }();
    return v;
}

// Contiguous rows of at least this many points take three passes of small tiles instead of two passes of 64..128 KiB
// tiles.  Measured on B200 (f64, 2 GiB of rows): 2^22: 6.9 -> 7.4 TFLOP/s, 2^24: 3.6 -> 8.0, 2^26: 2.2 -> 8.1 (the
// 2048..8192-point column tiles of the two-pass form have 16..32 B rows and one CTA per SM); 2^20 / 2^21 are faster
// in two passes (7.7 / 7.5 vs 6.9 / 7.2).
static int64_t three_level_min() {
    const int64_t v = [] {
        const char* e = knob_env("SFC_THREE_LEVEL_MIN");
        return e ? atoll(e) : (int64_t)1 << 22;
    }();
    return v;
}
// Bluestein with M >= this many points runs as five passes of small tiles (three-level M-point transforms) instead of
// three passes of 64 KiB tiles; 0 = never.  Measured on B200: parity identical (2e-15), but 32 x 1,000,003: 1.70 -> 2.06 ms
// and 32 x 3^13: 3.54 -> 4.00 ms — with chirp generation, twiddles and the fused transform pair the small-tile passes
// are SM-bound too (4.7 TB/s average), so the two extra passes cost more than they save: off.
static int64_t blue3_min() {
    const int64_t v = [] {
        const char* e = knob_env("SFC_BLUE3_MIN");
        return e ? atoll(e) : 0;
    }();
    return v;
}
static int row_fourstep_len() {
    const int v = [] {
        const char* e = knob_env("SFC_ROW_FOURSTEP");
        return e ? atoi(e) : 0;
    }();
    return v;
}

// the three-pass plan for large 2-D transforms was written after the round's GPU budget was spent: off until it has run
static bool fft2_tile2d_enabled() {
    const int v = [] {
        const char* e = knob_env("SFC_FFT2_TILE2D");
        return e ? atoi(e) : 0;
    }();
    return v != 0;
}

static int64_t scratch_budget_bytes() {
    const int64_t v = [] {
        const char* e = knob_env("SFC_WORK_MB");
        int64_t mb = e ? atoll(e) : 2048;
        if (mb < 1) mb = 1;
        return mb << 20;
    }();
    return v;
}

struct ArrayRef {
    int role;
    bool real;
    int64_t n;  // extent of the transform axis in this array
};

struct PlanBuilder {
    Plan& pl;
    int prec;
    size_t cs;  // complex element bytes
    size_t rs;  // real element bytes
    PlanError& err;
    int64_t dev_bytes = 0;
    int scatter_parts = 0;  // > 1: the next single-pass axis stores through the scatter table
    int64_t scatter_pitch = 0;  // element pitch of the split axis inside a destination block (0 = inner extent)
    // SFC_DESC_AUX_MUL: tables multiplied into the next axis on load (indexed by input position) and
    // on store (indexed by output position); power-of-two lengths only
    const void* aux_in = nullptr;
    const void* aux_out = nullptr;

    bool fail(int code, const std::string& m) {
        err = {code, m};
        return false;
    }

    void need_ms(size_t bytes) { pl.ms_bytes_ = std::max(pl.ms_bytes_, bytes); }
    void need_sa(size_t bytes) { pl.sa_bytes_ = std::max(pl.sa_bytes_, bytes); }

    static void set_io(IoDesc& d, int64_t bs, int64_t os, int64_t is, int64_t es, int64_t len, int64_t pes,
                       int64_t pls) {
        d.ptr = nullptr;
        d.batch_stride = bs;
        d.outer_stride = os;
        d.inner_stride = is;
        d.elem_stride = es;
        d.len = len;
        d.pos_es = pes;
        d.pos_ls = pls;
    }

    bool finish_tile(Step& s, int64_t nlanes, int64_t inner, int64_t nbatch, const char* what) {
        if (!s.k) return fail(SFC_ERR_PLAN, std::string("no kernel instantiation for ") + what);
        if (nlanes <= 0 || nlanes > 0xFFFFFFFFLL) return fail(SFC_ERR_VALUE, "too many lanes for one pass");
        s.p.nlanes = (uint32_t)nlanes;
        s.p.inner_count = (uint32_t)std::max<int64_t>(inner, 1);
        const int64_t tiles = (nlanes + s.k->TL - 1) / s.k->TL;
        s.p.tiles_per_batch = (uint32_t)tiles;
        s.nbatch = nbatch;
        if (!s.scatter) s.p.peer_shift = -1;
        s.p.tw = table_stage_tw(prec, s.k->L, err);
        if (!s.p.tw) return false;
        {
            // positions are e*pos_es + lane_outer*pos_ls with e < L and lane_outer < nlanes/inner
            const int64_t max_lo = (nlanes - 1) / std::max<int64_t>(inner, 1);
            const int64_t in_max = (int64_t)(s.k->L - 1) * s.p.in.pos_es + max_lo * s.p.in.pos_ls;
            const int64_t out_max = (int64_t)(s.k->L - 1) * s.p.out.pos_es + max_lo * s.p.out.pos_ls;
            if (in_max < s.p.in.len && s.p.ld_op != LD_C2R) s.p.flags |= F_IN_NOMASK;
            if (out_max < s.p.out.len && s.p.st_op != ST_R2C) s.p.flags |= F_OUT_NOMASK;
            // pick the predicate-free flavour when nothing needs masking
            const bool full_tiles = nlanes % s.k->TL == 0;
            int mode = 0;
            if (full_tiles && fast_enabled()) {
                if (s.p.st_op == ST_R2C) {
                    if (s.p.ld_op == LD_C && in_max < s.p.in.len && s.p.in.elem_stride == 1 &&
                        s.p.out.elem_stride == 1 && s.p.out.len >= s.k->L + 1 &&
                        (s.p.flags & ~(uint32_t)(F_IN_NOMASK | F_OUT_NOMASK)) == 0)
                        mode = 2;
                } else if (s.p.ld_op == LD_C2R) {
                    if (s.p.in.elem_stride == 1 && s.p.in.len >= s.k->L + 1 && out_max < s.p.out.len &&
                        s.p.st_op == ST_C && !(s.p.flags & F_ST_REAL))
                        mode = 3;
                } else if (!(s.p.flags & F_ST_REAL) &&
                           (s.p.ld_op == LD_C || s.p.ld_op == LD_C_MUL || s.p.ld_op == LD_SPLIT2)) {
                    // masks (zero padding on load, crop on store) are handled by a per-thread element
                    // limit in 32-bit arithmetic
                    const int64_t lim = (int64_t)1 << 31;
                    if (s.p.in.len < lim && s.p.out.len < lim && in_max < lim && out_max < lim && s.p.in.pos_es < lim &&
                        s.p.out.pos_es < lim && s.p.in.pos_es > 0 && s.p.out.pos_es > 0)
                        mode = 1;
                }
            }
            if (mode) {
                const KernelEntry* f = flavour_of(s.k, mode);
                if (f) s.k = f;
            }
            // short contiguous rows: coalesced staging of the whole tile through shared memory (F_STAGE_*)
            if (mode == 1 && s.k->mode == 1 && s.k->L >= 8 && s.k->L <= stage_io_max_len() && s.p.inner_count == 1) {
                if (s.p.map_in == MAP_ROW && s.p.ld_op == LD_C && (s.p.flags & F_IN_NOMASK) && s.p.in.elem_stride == 1 &&
                    s.p.in.outer_stride == s.k->L)
                    s.p.flags |= F_STAGE_IN;
                if (s.p.map_out == MAP_ROW && (s.p.flags & F_OUT_NOMASK) && s.p.out.elem_stride == 1 &&
                    s.p.out.outer_stride == s.k->L && !s.scatter)
                    s.p.flags |= F_STAGE_OUT;
            }
            // contiguous rows whose lanes are whole warps (512 x 4, 1024 x 2): one named-barrier group per lane, so the exchanges
            // of a row never wait for the other rows of the tile (the streaming proxy: CTA-wide barriers are what costs)
            // (f32, 16 KiB tiles, six CTAs per SM: 90.9 -> 86.0 % and 90.4 -> 84.8 %, so f64 only unless SFC_ROW_LANE_GROUPS=2)
            if (mode == 1 && s.k->mode == 1 && s.k->groups == 1 && !s.k->dbl && row_lane_groups() &&
                (prec == PREC_F64 || row_lane_groups() > 1) && s.p.map_in == MAP_ROW &&
                s.p.map_out == MAP_ROW && s.k->TL > 1 && !(s.p.flags & (F_STAGE_IN | F_STAGE_OUT)) && !s.scatter) {
                int cnt = 0;
                const KernelEntry* t = kernel_table(&cnt);
                for (int i = 0; i < cnt; ++i)
                    if (t[i].mode == 1 && t[i].groups == s.k->TL && t[i].prec == s.k->prec && t[i].L == s.k->L && t[i].TL == s.k->TL &&
                        t[i].dbl == 0 && t[i].E == s.k->E) {
                        s.k = &t[i];
                        break;
                    }
            }
            // group-pipelined flavour for contiguous rows (two thread groups per CTA, one TMA landing buffer between them)
            if (mode == 1 && s.k->mode == 1 && gpipe_enabled() && (s.p.flags & F_IN_NOMASK) &&
                (s.p.ld_op == LD_C || s.p.ld_op == LD_C_MUL) && prec == PREC_F64 && s.p.map_in == MAP_ROW &&
                s.p.in.elem_stride == 1 && tiles * nbatch >= gpipe_min_tiles()) {
                int cnt = 0;
                const KernelEntry* t = kernel_table(&cnt);
                for (int i = 0; i < cnt; ++i)
                    if (t[i].mode == 4 && t[i].groups == 2 && t[i].prec == s.k->prec && t[i].L == s.k->L && t[i].TL == s.k->TL &&
                        t[i].dbl == s.k->dbl) {
                        s.k = &t[i];
                        break;
                    }
            }
            // late-prefetch persistent flavour: contiguous complex rows, no masks; the next tile lands in the idle exchange buffer
            // (knob value 1: tiles up to 64 KiB; 2: the 128 KiB row tiles too, instead of the split-exchange pipelined flavour)
            const int plv = pipe_late_enabled();
            const bool narrow = s.k->TL * (int)cs <= 64 && s.k->L * s.k->TL * (int)cs == 65536;  // the 64 KiB tiles with <= 64 B segments
            if (mode == 1 && s.k->mode == 1 && plv && !s.k->dbl && (s.p.flags & F_IN_NOMASK) &&
                (plv == 2 || plv == 3 || s.k->L * s.k->TL * (int)cs <= 65536) &&
                (s.p.ld_op == LD_C || s.p.ld_op == LD_C_MUL) && !(s.p.flags & F_STAGE_IN) &&
                tiles * nbatch * std::max<int64_t>(s.batch_mult, 1) >= pipe_late_min_tiles()) {
                const bool rows_ok = s.p.map_in == MAP_ROW && s.p.in.elem_stride == 1 &&
                                     (plv <= 3 || (narrow && s.group >= 0 && s.k->TL >= 2));
                // strided tiles: TL adjacent lanes must be adjacent in memory -> one tensor-map box per <= 256 elements
                int tm = 0;
                if (s.p.map_in == MAP_COL && (plv == 3 || (plv == 4 && narrow && s.group >= 0)) && s.k->TL <= 128) {
                    if ((s.p.inner_count == 1 && s.p.in.outer_stride == 1) ||
                        (s.p.in.inner_stride == 1 && s.p.in.outer_stride == (int64_t)s.p.inner_count))
                        tm = 1;
                    else if (s.p.in.inner_stride == 1 && (int64_t)s.p.inner_count % s.k->TL == 0)
                        tm = 2;
                }
                if (rows_ok || tm) {
                    const KernelEntry* f = flavour_of(s.k, 9);
                    if (f) {
                        s.k = f;
                        if (!rows_ok) {
                            s.tmap = tm;
                            s.p.flags |= F_TMAP_IN;
                        }
                    }
                }
            }
            // persistent TMA-pipelined flavour: unmasked complex loads of tiles whose TL lanes are adjacent in
            // memory (16-byte aligned bulk copies), and enough tiles for every resident CTA to pipeline a few
            // 128 KiB row tiles (one CTA per SM: nothing else overlaps its loads) always take it when they can
            const bool big_row = s.k->L * s.k->TL * (int)cs > 65536 && s.p.map_in == MAP_ROW && pipe_big_enabled();
            if (mode == 1 && s.k->mode == 1 && (pipe_enabled() || (big_row && plv != 2 && plv != 3)) && (s.p.flags & F_IN_NOMASK) &&
                (s.p.ld_op == LD_C || s.p.ld_op == LD_C_MUL) && prec == PREC_F64) {
                const bool rows_ok = s.p.map_in == MAP_ROW && s.p.in.elem_stride == 1;
                const bool cols_ok = pipe_enabled() == 1 && s.p.map_in == MAP_COL &&
                                     ((s.p.in.inner_stride == 1 && (int64_t)s.p.inner_count % s.k->TL == 0) ||
                                      (s.p.inner_count == 1 && s.p.in.outer_stride == 1));
                const int64_t total_tiles = tiles * nbatch;
                if ((rows_ok || cols_ok) && total_tiles >= (big_row ? 2 * 148 : pipe_min_tiles())) {
                    const KernelEntry* f = flavour_of(s.k, 4);
                    if (f) s.k = f;
                }
            }
        }
        if (s.p.flags & F_CHIRP_GEN) {
            // q = exp(-i*pi*2*D^2/N) for the per-thread position step D = (L/E) * pos_es, phase reduced exactly
            const int64_t tpl = s.k->L / s.k->E;
            auto qfor = [&](int64_t pes, double* q) {
                const unsigned __int128 D = (unsigned __int128)(tpl * pes);
                const uint64_t r = (uint64_t)((2 * D * D) % (unsigned __int128)s.p.chirp_mod);
                long double c, sn;
                unit_root((long double)r, (long double)s.p.chirp_mod, c, sn);
                q[0] = (double)c;
                q[1] = (double)sn;
            };
            qfor(s.p.in.pos_es, s.p.chirp_q_in);
            qfor(s.p.out.pos_es, s.p.chirp_q_out);
        }
        char buf[256];
        snprintf(buf, sizeof buf, "%s: tile L=%d TL=%d%s%s threads=%d smem=%zu lanes=%lld batches=%lld map=%s->%s", what,
                 s.k->L, s.k->TL, s.k->dbl ? " fwd*tab*inv" : "",
                 s.k->groups > 1 ? (s.k->mode == 4 ? " group-pipelined" : (s.k->groups == s.k->TL && s.k->mode ? " fast lane-groups" : (s.k->mode ? " fast 2-groups" : " generic 2-groups")))
                                  : (s.k->mode == 0 ? " generic" : (s.k->mode == 1 ? " fast" : (s.k->mode == 2 ? " fast-r2c" : (s.k->mode == 3 ? " fast-c2r" : (s.k->mode == 10 ? " radix-9/3" : (s.k->mode == 9 ? " late-prefetch" : (s.k->groups == 2 ? " group-pipelined" : " pipelined"))))))), s.k->threads, s.k->smem, (long long)nlanes,
                 (long long)nbatch, s.p.map_in == MAP_ROW ? "row" : "col", s.p.map_out == MAP_ROW ? "row" : "col");
        s.desc = buf;
        pl.steps_.push_back(s);
        return true;
    }

    int new_group(int64_t nbatch, int64_t bytes_per_batch) {
        Group g;
        g.nbatch = nbatch;
        g.chunk = std::max<int64_t>(1, scratch_budget_bytes() / std::max<int64_t>(bytes_per_batch, 1));
        g.chunk = std::min(g.chunk, nbatch);
        g.slice_bytes = (size_t)g.chunk * (size_t)bytes_per_batch;
        if (l2_chunk_bytes() > 0 && bytes_per_batch <= l2_max_batch_bytes() && nbatch > 1) {
            const int64_t c = std::max<int64_t>(1, l2_chunk_bytes() / bytes_per_batch);
            if (c < g.chunk) {
                g.chunk = c;
                g.slice_bytes = (size_t)c * (size_t)bytes_per_batch;
                const int64_t rounds = (nbatch + c - 1) / c;
                const int64_t fit = std::max<int64_t>(1, l2_total_bytes() / (int64_t)g.slice_bytes);
                g.ways = (int)std::min<int64_t>(std::min<int64_t>(l2_ways(), fit), rounds);
            }
        }
        pl.groups_.push_back(g);
        need_ms(g.slice_bytes * (size_t)g.ways);
        return (int)pl.groups_.size() - 1;
    }

    bool add_copy(ArrayRef src, ArrayRef dst, const std::vector<int64_t>& src_shape,
                  const std::vector<int64_t>& dst_shape, double scale, bool src_f64_override = true,
                  bool use_override = false) {
        Step s;
        s.kind = K_COPY;
        s.src = src.role;
        s.dst = dst.role;
        CopyParams& c = s.cp;
        c.ndim = (int)dst_shape.size();
        c.total = 1;
        for (int i = 0; i < c.ndim; ++i) {
            c.dst_shape[i] = dst_shape[i];
            c.src_shape[i] = src_shape[i];
            c.total *= dst_shape[i];
        }
        if (c.ndim == 0) {
            c.ndim = 1;
            c.dst_shape[0] = c.src_shape[0] = 1;
            c.total = 1;
        }
        c.src_complex = src.real ? 0 : 1;
        c.dst_complex = dst.real ? 0 : 1;
        c.src_f64 = use_override ? (src_f64_override ? 1 : 0) : (prec == PREC_F64);
        c.dst_f64 = (prec == PREC_F64);
        c.conj_src = 0;
        c.scale = scale;
        s.desc = "copy/pad/crop";
        dev_bytes += c.total * (int64_t)((src.real ? rs : cs) + (dst.real ? rs : cs));
        pl.steps_.push_back(s);
        return true;
    }

    // One 1-D transform of length n along the middle axis of [O][n][I] arrays.
    bool add_axis(int64_t n, int64_t O, int64_t I, ArrayRef src, ArrayRef dst, bool inverse, double scale,
                  bool store_real) {
        const int lmax = lmax_for(prec);
        const uint32_t fl_in = inverse ? F_CONJ_LD_PRE : 0;
        const uint32_t fl_out = (inverse ? F_CONJ_ST_POST : 0) | (store_real ? F_ST_REAL : 0);
        const size_t src_es = src.real ? rs : cs;
        const size_t dst_es = (dst.real || store_real) ? rs : cs;
        const bool col = I > 1;
        if (O <= 0 || I <= 0 || n <= 0) return fail(SFC_ERR_VALUE, "empty transform axis");

        if ((aux_in || aux_out) && (!is_pow2(n) || n == 1 || n > (int64_t)lmax * lmax))
            return fail(SFC_ERR_NOT_IMPLEMENTED, "fused table multiplies need a power-of-two transform length");
        if (n == 1) {
            // length-1 DFT is the identity
            std::vector<int64_t> ss{O, src.n, I}, ds{O, dst.n, I};
            ArrayRef d2 = dst;
            d2.real = dst.real || store_real;
            return add_copy(src, d2, ss, ds, scale);
        }

        // strided lanes need >= 64..128 B of adjacent lanes per element row; above this length a
        // whole-column tile no longer fits shared memory with that many lanes, so the axis is
        // split into two strided sub-passes (four-step) instead
        const int col_single_max = col_single_limit(prec);
        // Contiguous rows whose single tile would be 128 KiB (one CTA per SM): split every row over two
        // lanes of half the length with a radix-2 stage folded into the load (LD_SPLIT2)
        if (!col && is_pow2(n) && n <= lmax && n == row_split_len(prec) && !src.real && !store_real && src.n == n &&
            dst.n == n && O * 2 <= 0xFFFFFFFFLL) {
            Step s;
            const int L = (int)(n / 2);
            s.k = pick_kernel(prec, L, false, 0);
            if (s.k && (O * 2) % s.k->TL == 0) {
                s.src = src.role;
                s.dst = dst.role;
                s.src_esize = cs;
                s.dst_esize = cs;
                // lane = 2*row + parity: both parities read the same row, parity picks the output interleave
                set_io(s.p.in, 0, n, 0, 1, n, 1, 0);
                set_io(s.p.out, 0, n, 1, 2, n, 2, 0);
                s.p.map_in = s.p.map_out = MAP_ROW;
                s.p.ld_op = LD_SPLIT2;
                s.p.st_op = ST_C;
                s.p.flags = fl_in | fl_out;
                s.p.scale = scale;
                s.p.rtw = table_rtw(prec, L, err);
                if (!s.p.rtw) return false;
                dev_bytes += O * n * 2 * (int64_t)cs;
                if (!finish_tile(s, O * 2, 2, 1, "single-pass rows, each row split over two half-length lanes")) return false;
                if (pl.steps_.back().k->mode == 1) return true;
                pl.steps_.pop_back();  // no fast flavour: fall through to the plain single tile
            }
        }
        // contiguous rows whose single tile would be 128 KiB (one CTA per SM, no overlap): optionally two small
        // passes instead (four-step), L2-blocked so that the intermediate never leaves the cache (SFC_ROW_FOURSTEP)
        const bool row_fs = !col && is_pow2(n) && row_fourstep_len() > 0 && n >= row_fourstep_len() && !src.real &&
                            !store_real && !aux_in && !aux_out && scatter_parts <= 1;
        if (is_pow2(n) && n <= lmax && !(col && n > col_single_max) && !row_fs) {
            Step s;
            // measured (1024^3 f64): with element rows <= 16 KiB apart the 64 KiB two-per-SM tile wins
            // (74 % vs 61 %); with multi-MiB strides (TLB-bound) the wide 128-byte-row tile wins (59 % vs 44 %)
            // a pass whose store goes over NVLink (split-axis scatter) is link-bound, and the link likes long segments
            // (1024^3 on 8 GPUs: 64 B stores reach 426 GB/s, 128 B stores 667 GB/s): widest tile there
            if (col && scatter_parts > 1 && (O * I) % 8 == 0)
                s.k = pick_kernel(prec, (int)n, true, 0);
            else if (col && (int64_t)I * (int64_t)cs <= (256 << 10))
                s.k = pick_kernel_two_per_sm(prec, (int)n, 0);
            else
                s.k = pick_kernel(prec, (int)n, col, 0);
            s.src = src.role;
            s.dst = dst.role;
            s.src_esize = src_es;
            s.dst_esize = dst_es;
            set_io(s.p.in, 0, src.n * I, 1, I, std::min(n, src.n), 1, 0);
            set_io(s.p.out, 0, dst.n * I, 1, I, std::min(n, dst.n), 1, 0);
            s.p.map_in = s.p.map_out = col ? MAP_COL : MAP_ROW;
            s.p.ld_op = src.real ? LD_R : LD_C;
            s.p.st_op = ST_C;
            if (aux_in) {
                s.p.ld_op = src.real ? LD_R_MUL : LD_C_MUL;
                s.p.aux_in = aux_in;
            }
            if (aux_out) {
                s.p.st_op = ST_MUL;
                s.p.aux_out = aux_out;
            }
            s.p.flags = fl_in | fl_out;
            s.p.scale = scale;
            dev_bytes += O * I * (std::min(n, src.n) * (int64_t)src_es + std::min(n, dst.n) * (int64_t)dst_es);
            if (scatter_parts > 1) {
                const int64_t blk = n / scatter_parts;
                if (n % scatter_parts || !is_pow2(blk) || dst.n != n || store_real)
                    return fail(SFC_ERR_VALUE, "scatter needs the split axis length / parts to be a power of two");
                s.scatter = true;
                s.p.peer_shift = ilog2_64(blk);
                const int64_t pitch = scatter_pitch > 0 ? scatter_pitch : I;
                if (pitch < I) return fail(SFC_ERR_VALUE, "scatter_pitch must be at least the inner extent");
                set_io(s.p.out, 0, blk * pitch, 1, pitch, n, 1, 0);
                if (!finish_tile(s, O * I, I, 1, "single-pass axis + split-axis scatter store")) return false;
                if (pl.steps_.back().k->mode != 1)
                    return fail(SFC_ERR_NOT_IMPLEMENTED, "scatter store needs full tiles (lanes % tile lanes == 0)");
                return true;
            }
            return finish_tile(s, O * I, I, 1, "single-pass axis");
        }

        if (is_pow2(n)) {
            // four-step (Bailey): n = L1*L2, columns then rows, through the work area [O][n][I]
            const int lg = ilog2_64(n);
            const bool can3 = !(col || src.real || store_real || src.n != n || dst.n != n || aux_in || aux_out || scatter_parts > 1 ||
                                lg > 3 * ilog2_64(lmax) || O > 0x7FFFFFFF / 8192);
            if (n > (int64_t)lmax * lmax && !can3)
                return fail(SFC_ERR_NOT_IMPLEMENTED, "transform length above lmax^2 (only contiguous complex rows go three levels deep)");
            if (can3 && (n > (int64_t)lmax * lmax || (prec == PREC_F64 && three_level_min() > 0 && n >= three_level_min())))
                return add_three_level(n, O, src, dst, fl_in, fl_out, scale);
            int64_t L1 = (int64_t)1 << (lg / 2);
            int64_t L2 = n / L1;
            if (L2 > lmax) {
                L2 = lmax;
                L1 = n / L2;
            }
            const void *lo, *hi;
            int sh;
            if (!table_fourstep(prec, n, &lo, &hi, &sh, err)) return false;
            const int g = new_group(O, n * I * (int64_t)cs);
            Step a;
            a.k = pick_kernel_two_per_sm(prec, (int)L1, 0);
            a.src = src.role;
            a.dst = R_MS;
            a.src_esize = src_es;
            a.dst_esize = cs;
            a.group = g;
            set_io(a.p.in, src.n * I, I, 1, L2 * I, std::min(n, src.n), L2, 1);
            set_io(a.p.out, n * I, I, 1, L2 * I, n, L2, 1);
            a.p.map_in = a.p.map_out = MAP_COL;
            a.p.ld_op = src.real ? LD_R : LD_C;
            if (aux_in) {
                a.p.ld_op = src.real ? LD_R_MUL : LD_C_MUL;
                a.p.aux_in = aux_in;
            }
            a.p.st_op = ST_TW;
            a.p.tw_lo = lo;
            a.p.tw_hi = hi;
            a.p.tw_shift = sh;
            a.p.flags = fl_in;
            a.p.scale = 1.0;
            if (!finish_tile(a, L2 * I, I, O, "four-step pass A (columns + twiddle)")) return false;
            Step b;
            b.k = pick_kernel_two_per_sm(prec, (int)L2, 0);
            b.src = R_MS;
            b.dst = dst.role;
            b.src_esize = cs;
            b.dst_esize = dst_es;
            b.group = g;
            set_io(b.p.in, n * I, L2 * I, 1, I, L2, 1, 0);
            set_io(b.p.out, dst.n * I, I, 1, L1 * I, std::min(n, dst.n), L1, 1);
            b.p.map_in = col ? MAP_COL : MAP_ROW;
            b.p.map_out = MAP_COL;
            b.p.ld_op = LD_C;
            b.p.st_op = ST_C;
            if (aux_out) {
                b.p.st_op = ST_MUL;
                b.p.aux_out = aux_out;
            }
            b.p.flags = fl_out;
            b.p.scale = scale;
            dev_bytes += O * I * (std::min(n, src.n) * (int64_t)src_es + 2 * n * (int64_t)cs +
                                  std::min(n, dst.n) * (int64_t)dst_es);
            return finish_tile(b, L1 * I, I, O, "four-step pass B (rows, transposed store)");
        }

        // n = 3^k (rustfft: Radix3): power-of-three tiles, one pass up to 3^7 = 2187, two passes (four-step) up to 3^14
        const int k3 = ilog3_exact(n);
        if (radix3_enabled() && k3 >= 2 && k3 <= 14 && !src.real && !store_real && src.n == n && dst.n == n && !aux_in && !aux_out &&
            scatter_parts <= 1 && O * I <= 0x7FFFFFFFLL) {
            if (k3 <= 7) {
                Step s;
                s.k = pick_r3(prec, (int)n, col);
                s.src = src.role;
                s.dst = dst.role;
                s.src_esize = cs;
                s.dst_esize = cs;
                set_io(s.p.in, 0, n * I, 1, I, n, 1, 0);
                set_io(s.p.out, 0, n * I, 1, I, n, 1, 0);
                s.p.map_in = s.p.map_out = col ? MAP_COL : MAP_ROW;
                s.p.ld_op = LD_C;
                s.p.st_op = ST_C;
                s.p.flags = fl_in | fl_out;
                s.p.scale = scale;
                dev_bytes += O * I * 2 * n * (int64_t)cs;
                return finish_tile(s, O * I, I, 1, "single-pass axis, power-of-three tile");
            }
            int64_t L1 = 1;
            for (int i = 0; i < k3 / 2; ++i) L1 *= 3;
            const int64_t L2 = n / L1;
            const void *lo, *hi;
            int sh;
            if (!table_fourstep(prec, n, &lo, &hi, &sh, err)) return false;
            const int g = new_group(O, n * I * (int64_t)cs);
            Step a;
            a.k = pick_r3(prec, (int)L1, true);
            a.src = src.role;
            a.dst = R_MS;
            a.src_esize = cs;
            a.dst_esize = cs;
            a.group = g;
            set_io(a.p.in, n * I, I, 1, L2 * I, n, L2, 1);
            set_io(a.p.out, n * I, I, 1, L2 * I, n, L2, 1);
            a.p.map_in = a.p.map_out = MAP_COL;
            a.p.ld_op = LD_C;
            a.p.st_op = ST_TW;
            a.p.tw_lo = lo;
            a.p.tw_hi = hi;
            a.p.tw_shift = sh;
            a.p.flags = fl_in;
            a.p.scale = 1.0;
            if (!finish_tile(a, L2 * I, I, O, "four-step pass A (columns + twiddle), power-of-three tile")) return false;
            Step b;
            b.k = pick_r3(prec, (int)L2, true);  // its transposed store is the strided side
            b.src = R_MS;
            b.dst = dst.role;
            b.src_esize = cs;
            b.dst_esize = cs;
            b.group = g;
            set_io(b.p.in, n * I, L2 * I, 1, I, L2, 1, 0);
            set_io(b.p.out, n * I, I, 1, L1 * I, n, L1, 1);
            b.p.map_in = col ? MAP_COL : MAP_ROW;
            b.p.map_out = MAP_COL;
            b.p.ld_op = LD_C;
            b.p.st_op = ST_C;
            b.p.flags = fl_out;
            b.p.scale = scale;
            dev_bytes += O * I * 4 * n * (int64_t)cs;
            return finish_tile(b, L1 * I, I, O, "four-step pass B (rows, transposed store), power-of-three tile");
        }

        // Bluestein chirp-z over a padded power-of-two convolution of length M >= 2n-1
        const int64_t M = next_pow2(2 * n - 1);
        const bool gen = chirp_gen_enabled() && M > lmax && M <= ((int64_t)1 << 26);  // position products < 2^52
        const void* chirp = gen ? nullptr : table_chirp(prec, n, err);
        if (!gen && !chirp) return false;
        const void *cr_lo = nullptr, *cr_hi = nullptr;
        int cr_sh = 0;
        if (gen && !table_chirp_roots(n, &cr_lo, &cr_hi, &cr_sh, err)) return false;
        auto set_chirp_gen = [&](Step& st) {
            st.p.flags |= F_CHIRP_GEN;
            st.p.chirp_lo = cr_lo;
            st.p.chirp_hi = cr_hi;
            st.p.chirp_shift = cr_sh;
            st.p.chirp_mod = (uint64_t)(2 * n);
        };
        if (M <= lmax) {
            const void* bf = table_bluestein_b(prec, n, M, 0, 0, err);
            if (!bf) return false;
            Step s;
            s.k = pick_kernel(prec, (int)M, col, 1);
            s.src = src.role;
            s.dst = dst.role;
            s.src_esize = src_es;
            s.dst_esize = dst_es;
            set_io(s.p.in, 0, src.n * I, 1, I, std::min(n, src.n), 1, 0);
            set_io(s.p.out, 0, dst.n * I, 1, I, std::min(n, dst.n), 1, 0);
            s.p.map_in = s.p.map_out = col ? MAP_COL : MAP_ROW;
            s.p.ld_op = src.real ? LD_R_MUL : LD_C_MUL;
            s.p.aux_in = chirp;
            s.p.mid = bf;
            s.p.mid_es = 1;
            s.p.mid_ls = 0;
            s.p.st_op = ST_MUL;
            s.p.aux_out = chirp;
            s.p.flags = fl_in | fl_out;
            s.p.scale = scale;
            dev_bytes += O * I * (std::min(n, src.n) * (int64_t)src_es + std::min(n, dst.n) * (int64_t)dst_es);
            return finish_tile(s, O * I, I, 1, "Bluestein single-pass (chirp*FFT*B*IFFT*chirp)");
        }
        if (M > (int64_t)lmax * lmax) return fail(SFC_ERR_NOT_IMPLEMENTED, "Bluestein length above lmax^2");
        const int lg = ilog2_64(M);
        if (gen && !col && prec == PREC_F64 && blue3_min() > 0 && M >= blue3_min() && O <= 0x7FFFFFFF / 8192) {
            // Five passes of small tiles (three-level M-point transforms, the innermost forward / inverse pair fused):
            //   A (chirp, pad, outer columns, W_M) -> A2 (inner columns, W_MM) -> B (rows: FFT * B' * IFFT * conj W_MM)
            //   -> C2 (inverse inner columns) -> C (conj W_M on load, inverse outer columns, chirp, crop)
            const int l1 = lg / 3, l2 = (lg - l1) / 2, l3 = lg - l1 - l2;
            const int64_t L1 = (int64_t)1 << l1, L2 = (int64_t)1 << l2, L3 = (int64_t)1 << l3, MM = L2 * L3;
            const void* bf = table_bluestein_b(prec, n, M, L1, L2, err, L3);
            if (!bf) return false;
            const void *lo_n, *hi_n, *lo_m, *hi_m;
            int sh_n, sh_m;
            if (!table_fourstep(prec, M, &lo_n, &hi_n, &sh_n, err)) return false;
            if (!table_fourstep(prec, MM, &lo_m, &hi_m, &sh_m, err)) return false;
            const int g = new_group(O, M * (int64_t)cs);
            {
                Step a;
                a.k = pick_kernel_two_per_sm(prec, (int)L1, 0);
                a.src = src.role;
                a.dst = R_MS;
                a.src_esize = src_es;
                a.dst_esize = cs;
                a.group = g;
                set_io(a.p.in, src.n, 1, 1, MM, std::min(n, src.n), MM, 1);
                set_io(a.p.out, M, 1, 1, MM, M, MM, 1);
                a.p.map_in = a.p.map_out = MAP_COL;
                a.p.ld_op = src.real ? LD_R_MUL : LD_C_MUL;
                a.batch_fastest = true;
                a.p.st_op = ST_TW;
                a.p.tw_lo = lo_n;
                a.p.tw_hi = hi_n;
                a.p.tw_shift = sh_n;
                a.p.flags = fl_in;
                a.p.scale = 1.0;
                set_chirp_gen(a);
                if (!finish_tile(a, MM, 1, O, "Bluestein-3 pass A (chirp, pad, outer columns, twiddle)")) return false;
            }
            {
                Step a2;
                a2.k = pick_kernel_two_per_sm(prec, (int)L2, 0);
                a2.src = R_MS;
                a2.dst = R_MS;
                a2.src_esize = cs;
                a2.dst_esize = cs;
                a2.group = g;
                a2.batch_mult = L1;
                set_io(a2.p.in, MM, 1, 1, L3, MM, L3, 1);
                set_io(a2.p.out, MM, 1, 1, L3, MM, L3, 1);
                a2.p.map_in = a2.p.map_out = MAP_COL;
                a2.p.ld_op = LD_C;
                a2.p.st_op = ST_TW;
                a2.p.tw_lo = lo_m;
                a2.p.tw_hi = hi_m;
                a2.p.tw_shift = sh_m;
                a2.p.flags = 0;
                a2.p.scale = 1.0;
                if (!finish_tile(a2, L3, 1, O, "Bluestein-3 pass A2 (inner columns, twiddle)")) return false;
            }
            {
                Step b;
                b.k = pick_kernel(prec, (int)L3, false, 1);
                b.src = R_MS;
                b.dst = R_MS;
                b.src_esize = cs;
                b.dst_esize = cs;
                b.group = g;
                // lane = k2*L1 + k1: lane_outer = k2 (the twiddle index), lane_inner = k1
                set_io(b.p.in, M, L3, MM, 1, L3, 1, 0);
                set_io(b.p.out, M, L3, MM, 1, L3, 1, 0);
                b.p.map_in = b.p.map_out = MAP_ROW;
                b.p.ld_op = LD_C;
                b.p.mid = bf;
                b.p.mid_es = 1;
                b.p.mid_ls = L3;
                b.p.mid_is = MM;
                b.p.st_op = ST_TW;
                b.p.tw_lo = lo_m;
                b.p.tw_hi = hi_m;
                b.p.tw_shift = sh_m;
                b.p.flags = F_TW_CONJ;
                b.p.scale = 1.0;
                b.batch_fastest = true;
                if (!finish_tile(b, L1 * L2, L1, O, "Bluestein-3 pass B (rows: FFT * B * IFFT * conj twiddle)")) return false;
            }
            {
                Step c2;
                c2.k = pick_kernel_two_per_sm(prec, (int)L2, 0);
                c2.src = R_MS;
                c2.dst = R_MS;
                c2.src_esize = cs;
                c2.dst_esize = cs;
                c2.group = g;
                c2.batch_mult = L1;
                set_io(c2.p.in, MM, 1, 1, L3, MM, L3, 1);
                set_io(c2.p.out, MM, 1, 1, L3, MM, L3, 1);
                c2.p.map_in = c2.p.map_out = MAP_COL;
                c2.p.ld_op = LD_C;
                c2.p.st_op = ST_C;
                c2.p.flags = F_CONJ_LD_POST | F_CONJ_ST_PRE;
                c2.p.scale = 1.0;
                if (!finish_tile(c2, L3, 1, O, "Bluestein-3 pass C2 (inverse inner columns)")) return false;
            }
            {
                Step c;
                c.k = pick_kernel_two_per_sm(prec, (int)L1, 0);
                c.src = R_MS;
                c.dst = dst.role;
                c.src_esize = cs;
                c.dst_esize = dst_es;
                c.group = g;
                set_io(c.p.in, M, 1, 1, MM, M, MM, 1);
                set_io(c.p.out, dst.n, 1, 1, MM, std::min(n, dst.n), MM, 1);
                c.p.map_in = c.p.map_out = MAP_COL;
                c.p.ld_op = LD_C;
                c.p.ld_tw_lo = lo_n;
                c.p.ld_tw_hi = hi_n;
                c.p.ld_tw_shift = sh_n;
                c.p.st_op = ST_MUL;
                c.batch_fastest = true;
                c.p.flags = F_LD_TW | F_LD_TW_CONJ | F_CONJ_LD_POST | F_CONJ_ST_PRE | fl_out;
                c.p.scale = scale;
                set_chirp_gen(c);
                dev_bytes += O * (std::min(n, src.n) * (int64_t)src_es + 9 * M * (int64_t)cs + std::min(n, dst.n) * (int64_t)dst_es);
                return finish_tile(c, MM, 1, O, "Bluestein-3 pass C (conj twiddle, inverse outer columns, chirp, crop)");
            }
        }
        // column passes (A, C) like short lanes (two CTAs per SM with >= 64 B rows); the fused row
        // pass B takes whatever is left
        // the fused row pass B is fastest on 4096-point rows (one 64 KiB tile, no spills): L2 = 4096 whenever that
        // leaves >= 64-point columns (measured: M = 2^21 61.8 % with 512 x 4096 vs 59.1 % with 1024 x 2048)
        int64_t L1 = std::min<int64_t>((int64_t)1 << (lg / 2), blue_l1_cap());
        if (!knob_env("SFC_BLUE_L1") && prec == PREC_F64 && M / 4096 >= 64 && M / 4096 <= 1024) L1 = M / 4096;
        int64_t L2 = M / L1;
        if (L2 > lmax) {
            L2 = lmax;
            L1 = M / L2;
        }
        const void* bf = table_bluestein_b(prec, n, M, L1, L2, err);
        if (!bf) return false;
        const void *lo, *hi;
        int sh;
        if (!table_fourstep(prec, M, &lo, &hi, &sh, err)) return false;
        const int g = new_group(O, M * I * (int64_t)cs);
        {
            Step a;
            a.k = pick_kernel_two_per_sm(prec, (int)L1, 0);
            a.src = src.role;
            a.dst = R_MS;
            a.src_esize = src_es;
            a.dst_esize = cs;
            a.group = g;
            set_io(a.p.in, src.n * I, I, 1, L2 * I, std::min(n, src.n), L2, 1);
            set_io(a.p.out, M * I, I, 1, L2 * I, M, L2, 1);
            a.p.map_in = a.p.map_out = MAP_COL;
            a.p.ld_op = src.real ? LD_R_MUL : LD_C_MUL;
            a.p.aux_in = chirp;
            a.batch_fastest = true;
            a.p.st_op = ST_TW;
            a.p.tw_lo = lo;
            a.p.tw_hi = hi;
            a.p.tw_shift = sh;
            a.p.flags = fl_in;
            a.p.scale = 1.0;
            if (gen) set_chirp_gen(a);
            if (!finish_tile(a, L2 * I, I, O, "Bluestein pass A (chirp, pad, columns, twiddle)")) return false;
        }
        {
            Step b;
            b.k = pick_kernel(prec, (int)L2, col, 1);
            b.src = R_MS;
            b.dst = R_MS;
            b.src_esize = cs;
            b.dst_esize = cs;
            b.group = g;
            set_io(b.p.in, M * I, L2 * I, 1, I, L2, 1, 0);
            set_io(b.p.out, M * I, L2 * I, 1, I, L2, 1, 0);
            b.p.map_in = b.p.map_out = col ? MAP_COL : MAP_ROW;
            b.p.ld_op = LD_C;
            b.p.mid = bf;
            b.p.mid_es = 1;
            b.p.mid_ls = L2;
            b.p.st_op = ST_TW;
            b.p.tw_lo = lo;
            b.p.tw_hi = hi;
            b.p.tw_shift = sh;
            b.p.flags = F_TW_CONJ;
            b.p.scale = 1.0;
            b.batch_fastest = true;
            if (!finish_tile(b, L1 * I, I, O, "Bluestein pass B (rows: FFT * B * IFFT * conj twiddle)")) return false;
        }
        {
            Step c;
            c.k = pick_kernel_two_per_sm(prec, (int)L1, 0);
            c.src = R_MS;
            c.dst = dst.role;
            c.src_esize = cs;
            c.dst_esize = dst_es;
            c.group = g;
            set_io(c.p.in, M * I, I, 1, L2 * I, M, L2, 1);
            set_io(c.p.out, dst.n * I, I, 1, L2 * I, std::min(n, dst.n), L2, 1);
            c.p.map_in = c.p.map_out = MAP_COL;
            c.p.ld_op = LD_C;
            c.p.st_op = ST_MUL;
            c.p.aux_out = chirp;
            c.batch_fastest = true;
            c.p.flags = F_CONJ_LD_POST | F_CONJ_ST_PRE | fl_out;
            c.p.scale = scale;
            if (gen) set_chirp_gen(c);
            dev_bytes += O * I * (std::min(n, src.n) * (int64_t)src_es + 5 * M * (int64_t)cs +
                                  std::min(n, dst.n) * (int64_t)dst_es);
            return finish_tile(c, L2 * I, I, O, "Bluestein pass C (inverse columns, chirp, crop)");
        }
    }

    // n = L1*L2*L3 > lmax^2, contiguous rows (I == 1): three passes through the work area [O][n]
    //   A : L1-point columns at stride L2*L3, twiddle W_n^(k1*m)                     (m = n2*L3 + n3)
    //   B1: inside every row k1 of L2*L3 points, L2-point columns at stride L3, twiddle W_(L2*L3)^(k2*n3)
    //   B2: L3-point rows, stored transposed: X[k1 + L1*(k2 + L2*k3)]
    bool add_three_level(int64_t n, int64_t O, ArrayRef src, ArrayRef dst, uint32_t fl_in, uint32_t fl_out, double scale) {
        const int lg = ilog2_64(n);
        const int l1 = lg / 3, l2 = (lg - l1) / 2, l3 = lg - l1 - l2;
        const int64_t L1 = (int64_t)1 << l1, L2 = (int64_t)1 << l2, L3 = (int64_t)1 << l3, MM = L2 * L3;
        const void *lo_n, *hi_n, *lo_m, *hi_m;
        int sh_n, sh_m;
        if (!table_fourstep(prec, n, &lo_n, &hi_n, &sh_n, err)) return false;
        if (!table_fourstep(prec, MM, &lo_m, &hi_m, &sh_m, err)) return false;
        const int g = new_group(O, n * (int64_t)cs);
        {
            Step a;
            a.k = pick_kernel_two_per_sm(prec, (int)L1, 0);
            a.src = src.role;
            a.dst = R_MS;
            a.src_esize = cs;
            a.dst_esize = cs;
            a.group = g;
            set_io(a.p.in, n, 1, 1, MM, n, MM, 1);
            set_io(a.p.out, n, 1, 1, MM, n, MM, 1);
            a.p.map_in = a.p.map_out = MAP_COL;
            a.p.ld_op = LD_C;
            a.p.st_op = ST_TW;
            a.p.tw_lo = lo_n;
            a.p.tw_hi = hi_n;
            a.p.tw_shift = sh_n;
            a.p.flags = fl_in;
            a.p.scale = 1.0;
            if (!finish_tile(a, MM, 1, O, "three-level pass A (outer columns + twiddle)")) return false;
        }
        {
            Step b;
            b.k = pick_kernel_two_per_sm(prec, (int)L2, 0);
            b.src = R_MS;
            b.dst = R_MS;
            b.src_esize = cs;
            b.dst_esize = cs;
            b.group = g;
            b.batch_mult = L1;
            set_io(b.p.in, MM, 1, 1, L3, MM, L3, 1);
            set_io(b.p.out, MM, 1, 1, L3, MM, L3, 1);
            b.p.map_in = b.p.map_out = MAP_COL;
            b.p.ld_op = LD_C;
            b.p.st_op = ST_TW;
            b.p.tw_lo = lo_m;
            b.p.tw_hi = hi_m;
            b.p.tw_shift = sh_m;
            b.p.flags = 0;
            b.p.scale = 1.0;
            if (!finish_tile(b, L3, 1, O, "three-level pass B1 (inner columns + twiddle)")) return false;
        }
        {
            Step c;
            c.k = pick_kernel_two_per_sm(prec, (int)L3, 0);
            c.src = R_MS;
            c.dst = dst.role;
            c.src_esize = cs;
            c.dst_esize = cs;
            c.group = g;
            // lane = k2*L1 + k1 (k1 fastest: adjacent lanes are adjacent outputs)
            set_io(c.p.in, n, L3, MM, 1, n, 1, 0);
            set_io(c.p.out, n, L1, 1, L1 * L2, n, 1, 0);
            c.p.map_in = MAP_ROW;
            c.p.map_out = MAP_COL;
            c.p.ld_op = LD_C;
            c.p.st_op = ST_C;
            c.p.flags = fl_out;
            c.p.scale = scale;
            dev_bytes += O * 6 * n * (int64_t)cs;
            return finish_tile(c, L1 * L2, L1, O, "three-level pass B2 (rows, transposed store)");
        }
    }

    bool r2c_fast_ok(int64_t n, int64_t I) const {
        return I == 1 && is_pow2(n) && n >= 64 && n / 2 <= lmax_for(prec);
    }

    // rows of n reals -> rows of n/2+1 complex (rfft.rs:39-59)
    bool add_r2c(int64_t n, int64_t O, ArrayRef src, ArrayRef dst, double scale) {
        Step s;
        const int L = (int)(n / 2);
        s.k = pick_kernel(prec, L, false, 0);
        s.src = src.role;
        s.dst = dst.role;
        s.src_esize = cs;  // addressed as packed complex
        s.dst_esize = cs;
        set_io(s.p.in, 0, L, 1, 1, L, 1, 0);
        set_io(s.p.out, 0, dst.n, 1, 1, std::min<int64_t>(L + 1, dst.n), 1, 0);
        s.p.map_in = s.p.map_out = MAP_ROW;
        s.p.ld_op = LD_C;
        s.p.st_op = ST_R2C;
        s.p.flags = 0;
        s.p.scale = scale;
        s.p.rtw = table_rtw(prec, L, err);
        if (!s.p.rtw) return false;
        dev_bytes += O * (n * (int64_t)rs + std::min<int64_t>(L + 1, dst.n) * (int64_t)cs);
        return finish_tile(s, O, 1, 1, "real->complex fused pack + post-twiddle");
    }

    // lanes of n reals -> lanes of n reals along the middle axis of [O][n][I]: DCT/DST II or III (dct.rs:523-684,
    // dst.rs:484-592) on the packed n/2-point transform.  The I/O descriptors count REAL elements.
    bool add_dct2(int64_t n, int64_t O, int64_t I, ArrayRef src, ArrayRef dst, double scale, double scale_dc, bool type3,
                  bool sine) {
        Step s;
        const int L = (int)(n / 2);
        const bool col = I > 1;
        int cnt = 0;
        const KernelEntry* t = kernel_table(&cnt);
        s.k = nullptr;
        for (int i = 0; i < cnt; ++i)  // rows: the narrowest tile (more CTAs per SM); strided lanes: the widest
            if (t[i].prec == prec && t[i].L == L && t[i].mode == (type3 ? 6 : 5) &&
                (!s.k || (col ? t[i].TL > s.k->TL : t[i].TL < s.k->TL)))
                s.k = &t[i];
        const int64_t lanes = O * I;
        if (!s.k || lanes % s.k->TL != 0 || (col && I % s.k->TL != 0) || lanes > 0xFFFFFFFFLL)
            return fail(SFC_ERR_NOT_IMPLEMENTED, "no fused DCT kernel for this length / batch");
        if (col && (int64_t)s.k->TL * (int64_t)rs < 32)
            return fail(SFC_ERR_NOT_IMPLEMENTED, "column tiles of the fused DCT kernel would be narrower than a sector");
        s.src = src.role;
        s.dst = dst.role;
        s.src_esize = rs;
        s.dst_esize = rs;
        set_io(s.p.in, 0, n * I, 1, I, n, 1, 0);
        set_io(s.p.out, 0, n * I, 1, I, n, 1, 0);
        s.p.map_in = s.p.map_out = col ? MAP_COL : MAP_ROW;
        s.p.ld_op = LD_C;
        s.p.st_op = ST_C;
        s.p.flags = F_IN_NOMASK | F_OUT_NOMASK | (sine ? F_TRIG_SINE : 0);
        s.p.scale = scale;
        s.p.scale_dc = scale_dc;
        s.p.rtw = table_rtw(prec, L, err);
        if (!s.p.rtw) return false;
        s.p.aux_out = roots_table(TK_DCT2, prec, L + 1, 4.0L * (long double)n, 1.0L, n, err);  // exp(-i pi k / (2n)), k <= n/2
        if (!s.p.aux_out) return false;
        s.p.peer_shift = -1;
        s.p.tw = table_stage_tw(prec, L, err);
        if (!s.p.tw) return false;
        s.p.nlanes = (uint32_t)lanes;
        s.p.inner_count = (uint32_t)I;
        s.p.tiles_per_batch = (uint32_t)(lanes / s.k->TL);
        s.nbatch = 1;
        dev_bytes += lanes * n * 2 * (int64_t)rs;
        char buf[200];
        snprintf(buf, sizeof buf, "fused %s-%s %s (Makhoul packing): tile L=%d TL=%d threads=%d smem=%zu lanes=%lld",
                 sine ? "DST" : "DCT", type3 ? "III" : "II", col ? "columns" : "rows", s.k->L, s.k->TL, s.k->threads, s.k->smem,
                 (long long)lanes);
        s.desc = buf;
        pl.steps_.push_back(s);
        return true;
    }

    // 2-D transform of [R][C] complex in THREE passes of small tiles (emulated in
    // tools/fft2_three_pass_emulation.py): R = A*16, C = LB*256.  Experimental (SFC_FFT2_TILE2D=1).
    bool add_fft2_three_pass(int64_t R, int64_t C, double scale, bool inverse) {
        const int64_t LA = 16, A = R / LA, Cb = 256, LB = C / Cb, total = R * C;
        if (R % LA || C % Cb || !is_pow2(A) || !is_pow2(LB) || A > lmax_for(prec) || LA * LB > lmax_for(prec) || LB < 2)
            return fail(SFC_ERR_NOT_IMPLEMENTED, "three-pass 2-D plan: unsupported extents");
        need_ms((size_t)total * cs);
        need_sa((size_t)total * cs);
        // pass 1 = four-step pass A of axis 0 with L1 = A, L2 = 16: transforms over the high row digit, W_R^(k1*r_lo) on store
        const void *lo, *hi;
        int sh;
        if (!table_fourstep(prec, R, &lo, &hi, &sh, err)) return false;
        Step a;
        a.k = pick_kernel_two_per_sm(prec, (int)A, 0);
        a.src = R_IN;
        a.dst = R_MS;
        a.src_esize = a.dst_esize = cs;
        set_io(a.p.in, R * C, C, 1, LA * C, R, LA, 1);
        set_io(a.p.out, R * C, C, 1, LA * C, R, LA, 1);
        a.p.map_in = a.p.map_out = MAP_COL;
        a.p.ld_op = LD_C;
        a.p.st_op = ST_TW;
        a.p.tw_lo = lo;
        a.p.tw_hi = hi;
        a.p.tw_shift = sh;
        a.p.flags = inverse ? F_CONJ_LD_PRE : 0;  // the inverse is the forward plan on conjugated data (as everywhere else)
        a.p.scale = 1.0;
        if (!finish_tile(a, LA * C, C, 1, "2-D three-pass: pass 1 (high row digit + twiddle)")) return false;
        // pass 2 = the 16 x LB two-dimensional tile: lanes (k1, c_rest), elements e = LB*r_lo + c_hi at stride Cb
        Step m;
        int cnt = 0;
        const KernelEntry* t = kernel_table(&cnt);
        for (int i = 0; i < cnt; ++i)
            if (t[i].prec == prec && t[i].L == (int)(LA * LB) && t[i].mode == 8 && (!m.k || t[i].TL > m.k->TL)) m.k = &t[i];
        if (!m.k || Cb % m.k->TL != 0) return fail(SFC_ERR_NOT_IMPLEMENTED, "three-pass 2-D plan: no 2-D tile kernel for this shape");
        m.src = R_MS;
        m.dst = R_SA;
        m.src_esize = m.dst_esize = cs;
        set_io(m.p.in, 0, LA * C, 1, Cb, LA * LB, 1, 0);
        set_io(m.p.out, 0, C, 1, Cb, LA * LB, 1, 0);
        m.p.mid_es = A * C;  // k2  -> row k1 + A*k2
        m.p.mid_ls = Cb;     // kc1 -> column kc1*Cb + c_rest
        m.p.aux_out = table_stage_tw(prec, (int)C, err);  // W_C^j, j < C: the pass-A twiddle of axis 1, W_C^(kc1*c_rest)
        if (!m.p.aux_out) return false;
        m.p.map_in = m.p.map_out = MAP_COL;
        m.p.ld_op = LD_C;
        m.p.st_op = ST_C;
        m.p.flags = F_IN_NOMASK | F_OUT_NOMASK;
        m.p.scale = 1.0;
        m.p.peer_shift = -1;
        m.p.tw = table_stage_tw(prec, (int)(LA * LB), err);
        if (!m.p.tw) return false;
        m.p.nlanes = (uint32_t)(A * Cb);
        m.p.inner_count = (uint32_t)Cb;
        m.p.tiles_per_batch = (uint32_t)(A * Cb / m.k->TL);
        m.nbatch = 1;
        char buf[200];
        snprintf(buf, sizeof buf, "2-D three-pass: pass 2 (16 x %lld two-dimensional tile): tile L=%d TL=%d threads=%d smem=%zu lanes=%lld",
                 (long long)LB, m.k->L, m.k->TL, m.k->threads, m.k->smem, (long long)(A * Cb));
        m.desc = buf;
        pl.steps_.push_back(m);
        // pass 3 = four-step pass B of axis 1 with L1 = LB, L2 = 256: contiguous row segments, output column kc1 + LB*kc2
        Step b;
        // the narrowest tile that still stores 128-byte segments (8 lanes): 32 KiB, four CTAs per SM
        b.k = nullptr;
        for (int i = 0; i < cnt; ++i)
            if (t[i].prec == prec && t[i].L == (int)Cb && t[i].mode == 0 && t[i].groups == 1 && !t[i].dbl && t[i].E == 16 &&
                t[i].TL >= 8 && LB % t[i].TL == 0 && (!b.k || t[i].TL < b.k->TL))
                b.k = &t[i];
        if (!b.k) b.k = pick_kernel_two_per_sm(prec, (int)Cb, 0);
        b.src = R_SA;
        b.dst = R_OUT;
        b.src_esize = b.dst_esize = cs;
        set_io(b.p.in, C, Cb, 1, 1, Cb, 1, 0);
        set_io(b.p.out, C, 1, 1, LB, C, LB, 1);
        b.p.map_in = MAP_ROW;
        b.p.map_out = MAP_COL;
        b.p.ld_op = LD_C;
        b.p.st_op = ST_C;
        b.p.flags = inverse ? F_CONJ_ST_POST : 0;
        b.p.scale = scale;
        dev_bytes += 6 * total * (int64_t)cs;
        return finish_tile(b, LB, 1, R, "2-D three-pass: pass 3 (low column digit, transposed store)");
    }

    // DCT-IV / DST-IV of n reals per lane in one kernel (TM_FAST_DCT4): rows (I == 1) or a strided axis with I adjacent
    // lanes, addressed exactly like add_dct2.  Experimental, see fft_tile.cuh.
    bool add_dct4(int64_t n, int64_t O, int64_t I, ArrayRef src, ArrayRef dst, double scale, bool sine) {
        Step s;
        const int L = (int)(n / 2);
        const bool col = I > 1;
        int cnt = 0;
        const KernelEntry* t = kernel_table(&cnt);
        s.k = nullptr;
        for (int i = 0; i < cnt; ++i)  // rows: the narrowest tile (more CTAs per SM); strided lanes: the widest
            if (t[i].prec == prec && t[i].L == L && t[i].mode == 7 && (!s.k || (col ? t[i].TL > s.k->TL : t[i].TL < s.k->TL)))
                s.k = &t[i];
        const int64_t lanes = O * I;
        if (!s.k || lanes % s.k->TL != 0 || (col && I % s.k->TL != 0) || lanes > 0xFFFFFFFFLL)
            return fail(SFC_ERR_NOT_IMPLEMENTED, "no fused DCT-IV kernel for this length / batch");
        if (col && (int64_t)s.k->TL * (int64_t)rs < 32)
            return fail(SFC_ERR_NOT_IMPLEMENTED, "column tiles of the fused DCT-IV kernel would be narrower than a sector");
        s.src = src.role;
        s.dst = dst.role;
        s.src_esize = rs;
        s.dst_esize = rs;
        set_io(s.p.in, 0, n * I, 1, I, n, 1, 0);
        set_io(s.p.out, 0, n * I, 1, I, n, 1, 0);
        s.p.map_in = s.p.map_out = col ? MAP_COL : MAP_ROW;
        s.p.ld_op = LD_C;
        s.p.st_op = ST_C;
        s.p.flags = F_IN_NOMASK | F_OUT_NOMASK | (sine ? F_TRIG_SINE : 0);
        s.p.scale = scale;
        s.p.aux_in = table_dct4_pre(prec, n, err);
        if (!s.p.aux_in) return false;
        s.p.aux_out = roots_table(TK_DCT4_POST, prec, L, 2.0L * (long double)n, 1.0L, n, err);  // exp(-i pi k / n), k < n/2
        if (!s.p.aux_out) return false;
        s.p.peer_shift = -1;
        s.p.tw = table_stage_tw(prec, L, err);
        if (!s.p.tw) return false;
        s.p.nlanes = (uint32_t)lanes;
        s.p.inner_count = (uint32_t)I;
        s.p.tiles_per_batch = (uint32_t)(lanes / s.k->TL);
        s.nbatch = 1;
        dev_bytes += lanes * n * 2 * (int64_t)rs;
        char buf[200];
        snprintf(buf, sizeof buf, "fused %s-IV %s (half-length complex transform): tile L=%d TL=%d threads=%d smem=%zu lanes=%lld",
                 sine ? "DST" : "DCT", col ? "columns" : "rows", s.k->L, s.k->TL, s.k->threads, s.k->smem, (long long)lanes);
        s.desc = buf;
        pl.steps_.push_back(s);
        return true;
    }

    // rows of src.n (<= n/2+1 used) complex -> rows of n reals (rfft.rs:92-178)
    bool add_c2r(int64_t n, int64_t O, ArrayRef src, ArrayRef dst, double scale) {
        Step s;
        const int L = (int)(n / 2);
        s.k = pick_kernel(prec, L, false, 0);
        s.src = src.role;
        s.dst = dst.role;
        s.src_esize = cs;
        s.dst_esize = cs;  // addressed as packed complex
        set_io(s.p.in, 0, src.n, 1, 1, std::min<int64_t>(L + 1, src.n), 1, 0);
        set_io(s.p.out, 0, L, 1, 1, L, 1, 0);
        s.p.map_in = s.p.map_out = MAP_ROW;
        s.p.ld_op = LD_C2R;
        s.p.st_op = ST_C;
        s.p.flags = F_CONJ_LD_POST | F_CONJ_ST_PRE;
        s.p.scale = scale;
        s.p.rtw = table_rtw(prec, L, err);
        if (!s.p.rtw) return false;
        dev_bytes += O * (std::min<int64_t>(L + 1, src.n) * (int64_t)cs + n * (int64_t)rs);
        return finish_tile(s, O, 1, 1, "complex->real fused pre-twiddle + unpack");
    }
};

// ---------------------------------------------------------------- Plan::create

static int64_t prod(const std::vector<int64_t>& v, size_t a, size_t b) {
    int64_t p = 1;
    for (size_t i = a; i < b; ++i) p *= v[i];
    return p;
}

std::shared_ptr<Plan> Plan::create(const sfc_desc& d, PlanError& err) {
    if (d.ndim < 1 || d.ndim > SFC_MAX_DIMS) {
        err = {SFC_ERR_VALUE, "ndim must be in 1..8"};
        return nullptr;
    }
    if (d.naxes < 0 || d.naxes > SFC_MAX_DIMS) {
        err = {SFC_ERR_VALUE, "naxes must be in 0..8"};
        return nullptr;
    }
    if (d.prec != SFC_PREC_F32 && d.prec != SFC_PREC_F64) {
        err = {SFC_ERR_VALUE, "unknown precision"};
        return nullptr;
    }
    std::vector<int64_t> shape(d.shape, d.shape + d.ndim);
    for (int64_t s : shape)
        if (s <= 0) {
            err = {SFC_ERR_VALUE, "shape entries must be positive"};
            return nullptr;
        }
    std::vector<int> axes(d.axes, d.axes + d.naxes);
    for (int a : axes)
        if (a < 0 || a >= d.ndim) {
            char b[96];
            snprintf(b, sizeof b, "Axis %d out of bounds for array of dimension %d", a, d.ndim);
            err = {SFC_ERR_VALUE, b};
            return nullptr;
        }
    int devcount = 0;
    if (cudaGetDeviceCount(&devcount) != cudaSuccess || devcount == 0) {
        cudaGetLastError();
        err = {SFC_ERR_BACKEND, "no CUDA device available (this library has no CPU fallback)"};
        return nullptr;
    }

    std::shared_ptr<Plan> sp(new Plan());
    Plan& pl = *sp;
    pl.desc = d;
    pl.device = cur_device();
    const int prec = d.prec == SFC_PREC_F64 ? PREC_F64 : PREC_F32;
    const size_t cs = prec == PREC_F64 ? 16 : 8, rs = cs / 2;
    PlanBuilder B{pl, prec, cs, rs, err};

    const int64_t total = prod(shape, 0, shape.size());
    double log2sum = 0;
    for (int a : axes) log2sum += std::log2((double)shape[a]);
    int64_t alg = 0;
    int passes = 0;
    auto axis_passes = [&](int64_t n, bool plain_complex = false) {
        if (n == 1) return 0;
        const int lmax = lmax_for(prec);
        if (is_pow2(n)) return n <= lmax ? 1 : 2;  // algorithmic passes; long rows may physically take three
        const int k3 = ilog3_exact(n);
        if (plain_complex && radix3_enabled() && k3 >= 2 && k3 <= 14) return k3 <= 7 ? 1 : 2;  // power-of-three tiles
        return next_pow2(2 * n - 1) <= lmax ? 1 : 4;
    };

    bool ok = true;
    if (d.kind == SFC_C2C) {
        pl.in_elems = pl.out_elems = total;
        pl.in_esize = pl.out_esize = cs;
        const bool inv = d.direction == SFC_INVERSE;
        const bool real_in = (d.flags & SFC_DESC_REAL_INPUT) != 0;
        const bool real_out = (d.flags & SFC_DESC_REAL_OUTPUT) != 0;
        if (real_in) pl.in_esize = rs;
        if (real_out) pl.out_esize = rs;
        int64_t ax_in = 0, ax_out = 0;  // custom extents of the (single) transformed axis
        if (d.flags & (SFC_DESC_AXIS_LEN | SFC_DESC_AUX_MUL | SFC_DESC_REAL_OUTPUT)) {
            if (axes.size() != 1 || d.scatter_parts > 1) {
                err = {SFC_ERR_VALUE, "axis_in_len / axis_out_len / aux tables / real output need exactly one axis"};
                return nullptr;
            }
            if (d.flags & SFC_DESC_AXIS_LEN) {
                ax_in = d.axis_in_len > 0 ? d.axis_in_len : shape[axes[0]];
                ax_out = d.axis_out_len > 0 ? d.axis_out_len : shape[axes[0]];
                pl.in_elems = total / shape[axes[0]] * ax_in;
                pl.out_elems = total / shape[axes[0]] * ax_out;
            }
            if (d.flags & SFC_DESC_AUX_MUL) {
                B.aux_in = d.aux_in;
                B.aux_out = d.aux_out;
            }
        }
        if (axes.empty()) {
            ok = B.add_copy({R_IN, real_in, 1}, {R_OUT, false, 1}, shape, shape, d.scale);
        }
        // experimental: one plan for both axes of a large forward 2-D transform
        const bool three_pass_2d = fft2_tile2d_enabled() && prec == PREC_F64 && shape.size() == 2 && axes.size() == 2 &&
                                   axes[0] != axes[1] && shape[0] >= 256 && is_pow2(shape[0]) && (shape[1] == 4096 || shape[1] == 8192 || shape[1] == 16384) &&
                                   d.flags == 0 &&
                                   d.scatter_parts <= 1;
        if (three_pass_2d) {
            ok = B.add_fft2_three_pass(shape[0], shape[1], d.scale, inv);
            passes += 2;
            alg += 2 * 2 * total * (int64_t)cs;
        }
        for (size_t i = 0; !three_pass_2d && ok && i < axes.size(); ++i) {
            const int a = axes[i];
            const int64_t n = shape[a], O = prod(shape, 0, a), I = prod(shape, a + 1, shape.size());
            const bool last = (i + 1 == axes.size());
            B.scatter_parts = (last && d.scatter_parts > 1) ? d.scatter_parts : 0;
            B.scatter_pitch = B.scatter_parts ? d.scatter_pitch : 0;
            if (B.scatter_parts > 16) {
                err = {SFC_ERR_VALUE, "scatter_parts must be <= 16"};
                return nullptr;
            }
            if (B.scatter_parts && (n > col_single_limit(prec) || !is_pow2(n))) {
                err = {SFC_ERR_NOT_IMPLEMENTED, "scatter store needs a single-pass power-of-two split axis"};
                return nullptr;
            }
            // with a scatter the intermediate passes must not touch the caller's outputs: use scratch
            const int mid_role = d.scatter_parts > 1 ? R_SA : R_OUT;
            if (d.scatter_parts > 1 && axes.size() > 1) B.need_sa((size_t)total * cs);
            ok = B.add_axis(n, O, I, {i == 0 ? R_IN : mid_role, i == 0 && real_in, ax_in ? ax_in : n},
                            {last ? R_OUT : mid_role, false, ax_out ? ax_out : n}, inv, last ? d.scale : 1.0,
                            last && real_out);
            B.aux_in = B.aux_out = nullptr;
            B.scatter_parts = 0;
            const bool plain = !(i == 0 && real_in) && !(last && real_out) && !(ax_in && ax_in != n) && !(ax_out && ax_out != n) &&
                               !d.aux_in && !d.aux_out && d.scatter_parts <= 1;
            const int ap = axis_passes(n, plain);
            passes += ap;
            if (!is_pow2(n) && ap == 4)
                alg += O * I * (2 * n + 6 * next_pow2(2 * n - 1)) * (int64_t)cs;
            else
                alg += (int64_t)ap * 2 * total * (int64_t)cs;
        }
        pl.info.nominal_flops = 5.0 * (double)total * log2sum;
    } else if (d.kind == SFC_R2C) {
        if (axes.empty()) {
            err = {SFC_ERR_VALUE, "real transform needs at least one axis"};
            return nullptr;
        }
        const int la = axes.back();
        if (d.flags & SFC_DESC_DCT4) {
            const int64_t n = shape[la];
            if (axes.size() != 1 || !is_pow2(n) || n < 128 || n / 2 > lmax_for(prec) || prec != PREC_F64) {
                err = {SFC_ERR_NOT_IMPLEMENTED, "fused DCT-IV needs ONE f64 axis of a power-of-two length in 128..16384"};
                return nullptr;
            }
            pl.in_elems = pl.out_elems = total;
            pl.in_esize = pl.out_esize = rs;
            if (!B.add_dct4(n, prod(shape, 0, la), prod(shape, la + 1, shape.size()), {R_IN, true, n}, {R_OUT, true, n}, d.scale,
                            (d.flags & SFC_DESC_TRIG_SINE) != 0))
                return nullptr;
            pl.info.in_bytes = pl.info.out_bytes = total * (int64_t)rs;
            pl.info.algorithmic_bytes = 2 * total * (int64_t)rs;
            pl.info.device_bytes = B.dev_bytes;
            pl.info.num_passes = 1;
            pl.info.num_launches = 1;
            pl.info.nominal_flops = 2.5 * (double)total * std::log2((double)n);
            return sp;
        }
        if (d.flags & (SFC_DESC_DCT2 | SFC_DESC_DCT3)) {
            const int64_t n = shape[la];
            if (axes.size() != 1 || !is_pow2(n) || n < 128 || n / 2 > lmax_for(prec) || prec != PREC_F64) {
                err = {SFC_ERR_NOT_IMPLEMENTED, "fused DCT needs ONE f64 axis of a power-of-two length in 128..16384"};
                return nullptr;
            }
            pl.in_elems = pl.out_elems = total;
            pl.in_esize = pl.out_esize = rs;
            double dc = d.scale_dc != 0.0 ? d.scale_dc : 1.0;
            if (d.flags & SFC_DESC_DCT2_ORTHO0) dc *= 0.70710678118654752440;
            if (!B.add_dct2(n, prod(shape, 0, la), prod(shape, la + 1, shape.size()), {R_IN, true, n}, {R_OUT, true, n}, d.scale, dc,
                            (d.flags & SFC_DESC_DCT3) != 0, (d.flags & SFC_DESC_TRIG_SINE) != 0))
                return nullptr;
            pl.info.in_bytes = pl.info.out_bytes = total * (int64_t)rs;
            pl.info.algorithmic_bytes = 2 * total * (int64_t)rs;
            pl.info.device_bytes = B.dev_bytes;
            pl.info.num_passes = 1;
            pl.info.num_launches = 1;
            pl.info.nominal_flops = 2.5 * (double)total * std::log2((double)n);
            return sp;
        }
        std::vector<int64_t> hshape = shape;
        hshape[la] = shape[la] / 2 + 1;
        pl.in_elems = total;
        pl.in_esize = rs;
        pl.out_elems = prod(hshape, 0, hshape.size());
        pl.out_esize = cs;
        const int dup = (int)std::count(axes.begin(), axes.end(), la);
        if (dup > 1) {
            // transform at full size in scratch, crop at the end (rfft.rs:484-523 literally)
            B.need_sa((size_t)total * cs);
            for (size_t i = 0; ok && i < axes.size(); ++i) {
                const int a = axes[i];
                const int64_t n = shape[a], O = prod(shape, 0, a), I = prod(shape, a + 1, shape.size());
                ok = B.add_axis(n, O, I, {i == 0 ? R_IN : R_SA, i == 0, n}, {R_SA, false, n}, false, 1.0, false);
                passes += axis_passes(n);
                alg += (int64_t)axis_passes(n) * 2 * total * (int64_t)cs;
            }
            if (ok) ok = B.add_copy({R_SA, false, 1}, {R_OUT, false, 1}, shape, hshape, d.scale);
        } else {
            const int64_t n = shape[la], O = prod(shape, 0, la), I = prod(shape, la + 1, shape.size());
            const bool only = axes.size() == 1;
            if (B.r2c_fast_ok(n, I))
                ok = B.add_r2c(n, O, {R_IN, true, n}, {R_OUT, false, n / 2 + 1}, only ? d.scale : 1.0);
            else
                ok = B.add_axis(n, O, I, {R_IN, true, n}, {R_OUT, false, n / 2 + 1}, false, only ? d.scale : 1.0,
                                false);
            passes += axis_passes(n);
            alg += (int64_t)std::max(axis_passes(n), 1) * (total * (int64_t)rs + pl.out_elems * (int64_t)cs);
            for (size_t i = 0; ok && i + 1 < axes.size(); ++i) {
                const int a = axes[i];
                const int64_t m = hshape[a], O2 = prod(hshape, 0, a), I2 = prod(hshape, a + 1, hshape.size());
                const bool last = (i + 2 == axes.size());
                ok = B.add_axis(m, O2, I2, {R_OUT, false, m}, {R_OUT, false, m}, false, last ? d.scale : 1.0, false);
                passes += axis_passes(m);
                alg += (int64_t)axis_passes(m) * 2 * pl.out_elems * (int64_t)cs;
            }
        }
        pl.info.nominal_flops = 2.5 * (double)total * log2sum;
    } else if (d.kind == SFC_C2R) {
        if (axes.empty()) {
            err = {SFC_ERR_VALUE, "real transform needs at least one axis"};
            return nullptr;
        }
        const int la = axes.back();
        std::vector<int64_t> xshape = shape;
        xshape[la] = shape[la] / 2 + 1;
        const bool custom_in = (d.flags & SFC_DESC_CUSTOM_IN_SHAPE) != 0;
        if (custom_in)
            for (int i = 0; i < d.ndim; ++i) {
                xshape[i] = d.in_shape[i];
                if (xshape[i] <= 0) {
                    err = {SFC_ERR_VALUE, "in_shape entries must be positive"};
                    return nullptr;
                }
            }
        for (int i = 0; i < d.ndim; ++i)
            if (xshape[i] > shape[i]) {
                err = {SFC_ERR_DIMENSION, "input extent exceeds the output shape (the reference indexes out of bounds here)"};
                return nullptr;
            }
        pl.in_elems = prod(xshape, 0, xshape.size());
        pl.in_esize = cs;
        pl.out_elems = total;
        pl.out_esize = rs;
        std::vector<int64_t> hshape = shape;
        hshape[la] = shape[la] / 2 + 1;
        const int dup = (int)std::count(axes.begin(), axes.end(), la);
        const int64_t n = shape[la], I = prod(shape, la + 1, shape.size());
        const bool fast = dup == 1 && xshape == hshape && B.r2c_fast_ok(n, I);
        if (fast) {
            const bool only = axes.size() == 1;
            if (!only) B.need_sa((size_t)pl.in_elems * cs);
            for (size_t i = 0; ok && i + 1 < axes.size(); ++i) {
                const int a = axes[i];
                const int64_t m = hshape[a], O2 = prod(hshape, 0, a), I2 = prod(hshape, a + 1, hshape.size());
                ok = B.add_axis(m, O2, I2, {i == 0 ? R_IN : R_SA, false, m}, {R_SA, false, m}, true, 1.0, false);
                passes += axis_passes(m);
                alg += (int64_t)axis_passes(m) * 2 * pl.in_elems * (int64_t)cs;
            }
            const int64_t O = prod(shape, 0, la);
            if (ok) ok = B.add_c2r(n, O, {only ? R_IN : R_SA, false, n / 2 + 1}, {R_OUT, true, n}, d.scale);
            passes += 1;
            alg += pl.in_elems * (int64_t)cs + total * (int64_t)rs;
        } else {
            // literal reference algorithm: Hermitian fill -> complex inverse over all axes -> real part
            B.need_sa((size_t)total * cs);
            Step h;
            h.kind = K_HERM;
            h.src = R_IN;
            h.dst = R_SA;
            h.hp.ndim = d.ndim;
            h.hp.naxes = d.naxes;
            h.hp.total = total;
            h.hp.src_complex = 1;
            h.hp.f64 = prec == PREC_F64;
            for (int i = 0; i < d.ndim; ++i) {
                h.hp.out_shape[i] = shape[i];
                h.hp.x_shape[i] = xshape[i];
            }
            for (int i = 0; i < d.naxes; ++i) h.hp.axes[i] = axes[i];
            h.desc = "Hermitian reconstruction (rfft.rs:733-901)";
            pl.steps_.push_back(h);
            B.dev_bytes += pl.in_elems * (int64_t)cs + total * (int64_t)cs;
            for (size_t i = 0; ok && i < axes.size(); ++i) {
                const int a = axes[i];
                const int64_t m = shape[a], O2 = prod(shape, 0, a), I2 = prod(shape, a + 1, shape.size());
                const bool last = (i + 1 == axes.size());
                ok = B.add_axis(m, O2, I2, {R_SA, false, m}, {last ? R_OUT : R_SA, false, m}, true,
                                last ? d.scale : 1.0, last);
                passes += axis_passes(m);
                alg += (int64_t)axis_passes(m) * 2 * total * (int64_t)cs;
            }
        }
        pl.info.nominal_flops = 2.5 * (double)total * log2sum;
    } else {
        err = {SFC_ERR_VALUE, "unknown transform kind"};
        return nullptr;
    }
    if (!ok) return nullptr;

    // scratch (sa_ / ms_) is allocated by the first execution: cached plans that are never run again hold no memory
    pl.info.in_bytes = pl.in_elems * (int64_t)pl.in_esize;
    pl.info.out_bytes = pl.out_elems * (int64_t)pl.out_esize;
    pl.info.scratch_bytes = (int64_t)(pl.sa_bytes_ + pl.ms_bytes_);
    pl.info.algorithmic_bytes = alg;
    pl.info.device_bytes = B.dev_bytes;
    pl.info.num_passes = passes;
    int launches = 0;
    for (const Step& s : pl.steps_) {
        if (s.group >= 0) {
            const Group& g = pl.groups_[s.group];
            launches += (int)((g.nbatch + g.chunk - 1) / g.chunk);
        } else
            launches += 1;
    }
    pl.info.num_launches = launches;
    return sp;
}

static size_t (*g_pressure_hook)(const Plan*) = nullptr;
void set_scratch_pressure_hook(size_t (*hook)(const Plan* except)) { g_pressure_hook = hook; }

cudaError_t alloc_with_relief(void** p, size_t bytes, const Plan* except) {
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaErrorMemoryAllocation && g_pressure_hook) {
        cudaGetLastError();
        if (g_pressure_hook(except) > 0) e = cudaMalloc(p, bytes);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        *p = nullptr;
    }
    return e;
}

bool Plan::ensure_scratch(std::string& es) {
    if (sa_bytes_ && !sa_ && alloc_with_relief(&sa_, sa_bytes_, this) != cudaSuccess) {
        es = "cudaMalloc failed for plan scratch";
        return false;
    }
    if (ms_bytes_ && !ms_ && alloc_with_relief(&ms_, ms_bytes_, this) != cudaSuccess) {
        es = "cudaMalloc failed for plan work area";
        return false;
    }
    if ((sa_bytes_ || ms_bytes_) && !busy_ev_ && cudaEventCreateWithFlags(&busy_ev_, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        es = "cudaEventCreate failed";
        return false;
    }
    return true;
}

int Plan::prepare(std::string& es) {
    std::lock_guard<std::mutex> lk(mu_);
    return ensure_scratch(es) ? 0 : (int)SFC_ERR_MEMORY;
}

size_t Plan::release_scratch() {
    std::unique_lock<std::mutex> lk(mu_, std::try_to_lock);
    if (!lk.owns_lock()) return 0;  // being enqueued right now: leave it alone
    const size_t freed = scratch_resident();
    if (!freed) return 0;
    int prev = 0;
    cudaGetDevice(&prev);
    if (prev != device) cudaSetDevice(device);
    if (busy_valid_) cudaEventSynchronize(busy_ev_);
    if (sa_) cudaFree(sa_);
    if (ms_) cudaFree(ms_);
    sa_ = ms_ = nullptr;
    busy_valid_ = false;
    if (prev != device) cudaSetDevice(prev);
    return freed;
}

Plan::~Plan() {
    if (sa_) cudaFree(sa_);
    if (ms_) cudaFree(ms_);
    if (busy_ev_) cudaEventDestroy(busy_ev_);
    for (cudaEvent_t e : side_done_) cudaEventDestroy(e);
    for (cudaStream_t s : side_) cudaStreamDestroy(s);
    if (fork_ev_) cudaEventDestroy(fork_ev_);
}

bool Plan::ensure_side_streams(int n, std::string& es) {
    if (!fork_ev_ && cudaEventCreateWithFlags(&fork_ev_, cudaEventDisableTiming) != cudaSuccess) {
        es = "cudaEventCreate failed";
        return false;
    }
    while ((int)side_.size() < n) {
        cudaStream_t s;
        cudaEvent_t e;
        if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
            es = "could not create the side streams of an L2-blocked plan";
            cudaGetLastError();
            return false;
        }
        side_.push_back(s);
        side_done_.push_back(e);
    }
    return true;
}

std::string Plan::describe() const {
    std::string s;
    char b[160];
    snprintf(b, sizeof b, "plan kind=%d prec=%s dir=%d ndim=%d passes=%d launches=%d scratch=%lld B\n", desc.kind,
             desc.prec == SFC_PREC_F64 ? "f64" : "f32", desc.direction, desc.ndim, info.num_passes,
             info.num_launches, (long long)info.scratch_bytes);
    s += b;
    int i = 0;
    for (const Step& st : steps_) {
        snprintf(b, sizeof b, "  step %d: ", i++);
        s += b;
        s += st.desc;
        if (st.group >= 0) {
            snprintf(b, sizeof b, " [group %d: %lld batches, %lld per round, %d rounds in flight]", st.group,
                     (long long)groups_[st.group].nbatch, (long long)groups_[st.group].chunk, groups_[st.group].ways);
            s += b;
        }
        s += "\n";
    }
    return s;
}

// ------------------------------------------------------------------ Plan::exec

#ifdef SFC_PHASE_TIMING
static unsigned long long* g_dbg = nullptr;
static int g_dbg_launch = 0;
static std::vector<std::string> g_dbg_names;
static unsigned long long* dbg_slot(const std::string& name) {
    if (!g_dbg) {
        cudaMalloc(&g_dbg, 256 * 16 * 8);
        cudaMemset(g_dbg, 0, 256 * 16 * 8);
    }
    if (g_dbg_launch >= 256) return nullptr;
    g_dbg_names.push_back(name);
    return g_dbg + 16 * (g_dbg_launch++);
}
extern "C" __attribute__((visibility("default"))) void sfc_debug_phase_dump(void) {
    cudaDeviceSynchronize();
    std::vector<unsigned long long> h(256 * 16);
    cudaMemcpy(h.data(), g_dbg, h.size() * 8, cudaMemcpyDeviceToHost);
    for (int l = 0; l < g_dbg_launch; ++l) {
        const unsigned long long* r = &h[16 * l];
        const double n = (double)std::max<unsigned long long>(r[14], 1);
        double tot = 0;
        for (int k = 0; k < 7; ++k) tot += (double)r[k];
        printf("launch %d ctas %llu cyc/cta %.0f :", l, r[14], tot / n);
        for (int k = 0; k < 7; ++k) printf(" p%d %.0f", k, (double)r[k] / n);
        printf("  | %s\n", g_dbg_names[l].substr(0, 60).c_str());
    }
    cudaMemset(g_dbg, 0, 256 * 16 * 8);
    g_dbg_launch = 0;
    g_dbg_names.clear();
}
#endif

// 4-D tiled tensor map over the input of a strided pass: [batch][outer][element of the axis][2 * contiguous lanes] in units of
// the real type; one box = [rows][TL lanes].  Encoded per launch (the base pointer is a launch argument); ~1 us on the host.
static bool encode_tile_map(PassParams& p, const KernelEntry* k, int mode, int64_t nbatch, std::string& es) {
#ifdef SFC_HOST_EMUL
    (void)p; (void)k; (void)mode; (void)nbatch;
    es = "tensor maps are not emulated";
    return false;
#else
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) {
            cudaGetLastError();
            return (EncodeFn) nullptr;
        }
        return (EncodeFn)f;
    }();
    if (!fn) {
        es = "cuTensorMapEncodeTiled is not available from this driver";
        return false;
    }
    const bool f64 = k->prec == PREC_F64;
    const uint64_t cs = f64 ? 16 : 8;
    const uint64_t C = mode == 1 ? p.nlanes : p.inner_count;
    const uint64_t NO = mode == 1 ? 1 : p.nlanes / p.inner_count;
    cuuint64_t dims[4] = {2 * C, (cuuint64_t)k->L, NO, (cuuint64_t)std::max<int64_t>(nbatch, 1)};
    cuuint64_t str[3];
    str[0] = (cuuint64_t)p.in.elem_stride * cs;
    str[1] = NO > 1 ? (cuuint64_t)p.in.outer_stride * cs : str[0] * dims[1];
    str[2] = dims[3] > 1 ? (cuuint64_t)p.in.batch_stride * cs : str[1] * dims[2];
    const cuuint32_t rows = (cuuint32_t)std::min(k->L, 256);
    cuuint32_t box[4] = {(cuuint32_t)(2 * k->TL), rows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUtensorMap m;
    const CUresult r = fn(&m, f64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, p.in.ptr, dims, str, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char b[160];
        snprintf(b, sizeof b, "cuTensorMapEncodeTiled failed (%d): dims %llu x %llu x %llu x %llu", (int)r, (unsigned long long)dims[0],
                 (unsigned long long)dims[1], (unsigned long long)dims[2], (unsigned long long)dims[3]);
        es = b;
        return false;
    }
    static_assert(sizeof(CUtensorMap) == sizeof(p.tmap_in), "tensor map size");
    memcpy(p.tmap_in, &m, sizeof m);
    p.tmap_box_rows = (int32_t)rows;
    p.tmap_split = mode == 2 ? 1 : 0;
    return true;
#endif
}

bool Plan::window_ok(int64_t row_lanes, int nwin) const {
    if (steps_.size() != 1 || nwin < 1 || row_lanes < 1) return false;
    const Step& s = steps_[0];
    if (s.kind != K_TILE || s.group >= 0 || s.nbatch != 1 || s.tmap || !s.k || s.p.nbatch_fast) return false;
    if (row_lanes % s.k->TL) return false;
    const int64_t row_tiles = row_lanes / s.k->TL;
    return (int64_t)s.p.nlanes % row_lanes == 0 && row_tiles % nwin == 0 && (int64_t)s.p.tiles_per_batch * s.k->TL == (int64_t)s.p.nlanes;
}

int Plan::exec(const void* d_in, void* d_out, cudaStream_t stream, std::string& es, void* const* scatter, int nscatter,
               const ExecWindow* win) {
    std::lock_guard<std::mutex> lk(mu_);
    if (win && (!window_ok(win->row_lanes, win->count) || win->index < 0 || win->index >= win->count)) {
        es = "this plan cannot be executed over a window of its lanes";
        return SFC_ERR_VALUE;
    }
    bool uses_scratch = sa_bytes_ || ms_bytes_;
    if (uses_scratch) {
        // stream capture (CUDA graphs): no allocation and no cross-stream event inside a capture — the plan must have run
        // once before, and replays of the graph are ordered by the stream they are launched on
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (stream && cudaStreamIsCapturing(stream, &cap) != cudaSuccess) cudaGetLastError();
        if (cap != cudaStreamCaptureStatusNone) {
            if ((sa_bytes_ && !sa_) || (ms_bytes_ && !ms_)) {
                es = "execute the plan once before capturing it into a CUDA graph (its scratch is allocated on first use)";
                return SFC_ERR_VALUE;
            }
            uses_scratch = false;
        } else {
            if (!ensure_scratch(es)) return SFC_ERR_MEMORY;
            // the previous execution may still be running on another stream over the same scratch: order after it
            if (busy_valid_ && last_stream_ != stream) cudaStreamWaitEvent(stream, busy_ev_, 0);
        }
    }
    struct Mark {  // record the "scratch busy until here" event on every exit path that launched something
        Plan* pl; cudaStream_t st; bool on;
        ~Mark() {
            if (!on) return;
            cudaEventRecord(pl->busy_ev_, st);
            pl->busy_valid_ = true;
            pl->last_stream_ = st;
        }
    } mark{this, stream, uses_scratch};
    auto base = [&](int role) -> char* {
        switch (role) {
            case R_IN: return (char*)const_cast<void*>(d_in);
            case R_OUT: return (char*)d_out;
            case R_SA: return (char*)sa_;
            default: return (char*)ms_;
        }
    };
    auto fail = [&](cudaError_t e, const Step& s) {
        es = std::string("kernel launch failed (") + s.desc + "): " + cudaGetErrorString(e);
        return (int)SFC_ERR_COMPUTATION;
    };
    size_t i = 0;
    while (i < steps_.size()) {
        Step& s = steps_[i];
        if (s.kind == K_COPY) {
            s.cp.src = base(s.src);
            s.cp.dst = base(s.dst);
            cudaError_t e = launch_nd_copy(s.cp, stream);
            if (e != cudaSuccess) return fail(e, s);
            ++i;
            continue;
        }
        if (s.kind == K_HERM) {
            s.hp.src = base(s.src);
            s.hp.dst = base(s.dst);
            cudaError_t e = launch_herm_fill(s.hp, stream);
            if (e != cudaSuccess) return fail(e, s);
            ++i;
            continue;
        }
        if (s.group < 0) {
            PassParams p = s.p;
            p.in.ptr = base(s.src);
            p.out.ptr = base(s.dst);
            if (s.scatter) {
                if (!scatter || nscatter != desc.scatter_parts) {
                    es = "this plan stores through a scatter table: use sfc_exec_device_scatter with scatter_parts pointers";
                    return SFC_ERR_VALUE;
                }
                for (int q = 0; q < nscatter; ++q) p.peer_out[q] = scatter[q];
            }
            uint64_t grid = (uint64_t)p.tiles_per_batch * (uint64_t)s.nbatch;
            if (win) {
                p.win_row_tiles = (uint32_t)(win->row_lanes / s.k->TL);
                p.win_len = p.win_row_tiles / (uint32_t)win->count;
                p.win_first = p.win_len * (uint32_t)win->index;
                grid /= (uint64_t)win->count;
            }
            if (grid == 0 || grid > 0x7FFFFFFFULL) {
                es = "grid too large";
                return SFC_ERR_VALUE;
            }
#ifdef SFC_PHASE_TIMING
            p.dbg = knob_env("SFC_PHASE_DBG") ? dbg_slot(s.desc) : nullptr;
#endif
            if (s.tmap && !encode_tile_map(p, s.k, s.tmap, s.nbatch, es)) return SFC_ERR_BACKEND;
            cudaError_t e = s.k->launch(p, (unsigned)grid, stream);
            if (e != cudaSuccess) return fail(e, s);
            ++i;
            continue;
        }
        // chunk-looped group: steps [i, j) share the work area
        size_t j = i;
        while (j < steps_.size() && steps_[j].group == s.group) ++j;
        const Group& g = groups_[s.group];
        const int ways = g.ways;
        if (ways > 1) {
            if (!ensure_side_streams(ways, es)) return SFC_ERR_BACKEND;
            cudaEventRecord(fork_ev_, stream);
            for (int w = 0; w < ways; ++w) cudaStreamWaitEvent(side_[w], fork_ev_, 0);
        }
        int64_t round = 0;
        for (int64_t b0 = 0; b0 < g.nbatch; b0 += g.chunk, ++round) {
            const int64_t nb = std::min(g.chunk, g.nbatch - b0);
            const int way = ways > 1 ? (int)(round % ways) : 0;
            cudaStream_t st = ways > 1 ? side_[way] : stream;
            for (size_t k = i; k < j; ++k) {
                Step& t = steps_[k];
                PassParams p = t.p;
                char* ib = base(t.src);
                char* ob = base(t.dst);
                if (t.src == R_MS) ib += (size_t)way * g.slice_bytes;
                if (t.dst == R_MS) ob += (size_t)way * g.slice_bytes;
                if (t.src != R_MS) ib += (size_t)b0 * (size_t)p.in.batch_stride * t.src_esize;
                if (t.dst != R_MS) ob += (size_t)b0 * (size_t)p.out.batch_stride * t.dst_esize;
                p.in.ptr = ib;
                p.out.ptr = ob;
                if (t.batch_fastest) {
                    p.nbatch_fast = (uint32_t)nb;
                    int sh = tile_group_log2();
                    while (sh > 0 && (p.tiles_per_batch & ((1u << sh) - 1u))) --sh;
                    p.tile_group_shift = (uint32_t)sh;
                }
                const uint64_t grid = (uint64_t)p.tiles_per_batch * (uint64_t)nb * (uint64_t)t.batch_mult;
                if (grid == 0 || grid > 0x7FFFFFFFULL) {
                    es = "grid too large";
                    return SFC_ERR_VALUE;
                }
#ifdef SFC_PHASE_TIMING
                p.dbg = knob_env("SFC_PHASE_DBG") ? dbg_slot(t.desc) : nullptr;
#endif
                if (t.tmap && !encode_tile_map(p, t.k, t.tmap, nb * t.batch_mult, es)) return SFC_ERR_BACKEND;
                cudaError_t e = t.k->launch(p, (unsigned)grid, st);
                if (e != cudaSuccess) return fail(e, t);
            }
        }
        if (ways > 1) {
            for (int w = 0; w < ways; ++w) {
                cudaEventRecord(side_done_[w], side_[w]);
                cudaStreamWaitEvent(stream, side_done_[w], 0);
            }
        }
        i = j;
    }
    return 0;
}

}  // namespace sfc
