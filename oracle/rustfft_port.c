/* rustfft_port.c — CPU ORACLE ENGINE + CPU BASELINE.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this; the product (libscirs2_fft_cuda.so) never links or calls it.
 *
 * The reference (scirs2-fft, cool-japan/scirs 0.1.0-alpha.6) delegates all butterflies to
 * the third-party crate rustfft ("6.4.0", default-features = false => scalar planner,
 * /root/reference/Cargo.toml:86), which is not vendored and cannot be built here (no
 * cargo/rustc).  This file restates, from rustfft's published design, the scalar algorithm
 * CLASSES its planner composes — it is not a line-by-line port:
 *   - power-of-two lengths: iterative radix-4 decimation in time over a digit-reversed
 *     copy, with a radix-2 first level when log2(n) is odd        (rustfft `Radix4`)
 *   - other smooth lengths: recursive mixed-radix Cooley-Tukey, n = p * m with p the
 *     smallest prime factor <= 31, naive p-point butterflies       (rustfft `MixedRadix`,
 *     `Radix3`, `Butterfly*`)
 *   - lengths with a prime factor > 31: Bluestein over an inner power-of-two length
 *     >= 2n-1                                                      (rustfft `BluesteinsAlgorithm`;
 *     rustfft prefers Rader's algorithm for primes whose n-1 is smooth — same DFT, not restated)
 * Twiddles are cos/sin(-2*pi*k/n) evaluated in f64, as rustfft's `twiddles::compute_twiddle`.
 * Forward sign -, inverse +, both unnormalised, exactly like `Fft::process`.
 *
 * The rfp_ref_* entry points restate the reference's CALL PATTERN around `process`
 * (per-call planning, Vec copies, per-lane gather/scatter) for the timed CPU baseline:
 *   rfp_ref_fft        scirs2-fft/src/fft/algorithms.rs:131-176
 *   rfp_ref_rfft_rows  loop of scirs2-fft/src/rfft.rs:39-59 over the rows of a batch
 *   rfp_ref_irfft_rows loop of scirs2-fft/src/rfft.rs:92-178
 *   rfp_ref_fftn       scirs2-fft/src/fft/algorithms.rs:576-706 (lanes gather/process/scatter)
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double re, im; } cpx;

static inline cpx cmul(cpx a, cpx b) { cpx r = {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; return r; }
static inline cpx cadd(cpx a, cpx b) { cpx r = {a.re + b.re, a.im + b.im}; return r; }
static inline cpx csub(cpx a, cpx b) { cpx r = {a.re - b.re, a.im - b.im}; return r; }

static cpx twiddle(int64_t k, int64_t n, int inverse) {
    const double a = -2.0 * M_PI * (double)k / (double)n;
    cpx w = {cos(a), sin(a)};
    if (inverse) w.im = -w.im;
    return w;
}

/* ------------------------------------------------------------------ plan */

enum { ALG_DFT1, ALG_RADIX4, ALG_MIXED, ALG_BLUESTEIN };

typedef struct rfp_plan {
    int alg;
    int64_t n;
    int inverse;
    /* radix4 */
    cpx* tw;          /* n/4*3 per level packed, or generic n twiddles */
    int64_t* rev;     /* digit-reversal permutation */
    /* mixed radix */
    int64_t p, m;
    struct rfp_plan* sub;  /* length-m plan */
    cpx* tw_pm;       /* W_n^(j*k), j<p, k<m */
    cpx* bfly;        /* W_p^(a*b) */
    /* bluestein */
    int64_t M;
    struct rfp_plan* inner_f;
    struct rfp_plan* inner_i;
    cpx* chirp;       /* exp(-+ i*pi*k^2/n) */
    cpx* bspec;       /* FFT_M(conj chirp wrapped) / M */
} rfp_plan;

void rfp_plan_free(rfp_plan* p);

static int is_pow2(int64_t n) { return n > 0 && (n & (n - 1)) == 0; }

static int64_t smallest_factor(int64_t n) {
    if (n % 2 == 0) return 2;
    for (int64_t f = 3; f * f <= n; f += 2)
        if (n % f == 0) return f;
    return n;
}

static int64_t largest_prime_factor(int64_t n) {
    int64_t best = 1;
    while (n > 1) {
        int64_t f = smallest_factor(n);
        best = f > best ? f : best;
        n /= f;
    }
    return best;
}

rfp_plan* rfp_plan_new(int64_t n, int inverse) {
    rfp_plan* p = (rfp_plan*)calloc(1, sizeof(rfp_plan));
    p->n = n;
    p->inverse = inverse;
    if (n <= 1) {
        p->alg = ALG_DFT1;
        return p;
    }
    if (is_pow2(n)) {
        p->alg = ALG_RADIX4;
        p->tw = (cpx*)malloc(sizeof(cpx) * (size_t)n);
        for (int64_t k = 0; k < n; ++k) p->tw[k] = twiddle(k, n, inverse);
        /* digit reversal: base-4 digits from the LSB side (outer levels), then the one
         * remaining bit (innermost radix-2 level) when log2(n) is odd */
        int lg = 0;
        while (((int64_t)1 << lg) < n) ++lg;
        p->rev = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
        for (int64_t i = 0; i < n; ++i) {
            int64_t x = i, r = 0;
            int bits = lg;
            while (bits >= 2) { r = (r << 2) | (x & 3); x >>= 2; bits -= 2; }
            if (bits == 1) r = (r << 1) | (x & 1);
            p->rev[i] = r;
        }
        return p;
    }
    if (largest_prime_factor(n) <= 31) {
        p->alg = ALG_MIXED;
        p->p = smallest_factor(n);
        p->m = n / p->p;
        p->sub = rfp_plan_new(p->m, inverse);
        p->tw_pm = (cpx*)malloc(sizeof(cpx) * (size_t)n);
        for (int64_t j = 0; j < p->p; ++j)
            for (int64_t k = 0; k < p->m; ++k) p->tw_pm[j * p->m + k] = twiddle(j * k, n, inverse);
        p->bfly = (cpx*)malloc(sizeof(cpx) * (size_t)(p->p * p->p));
        for (int64_t a = 0; a < p->p; ++a)
            for (int64_t b = 0; b < p->p; ++b) p->bfly[a * p->p + b] = twiddle((a * b) % p->p, p->p, inverse);
        return p;
    }
    p->alg = ALG_BLUESTEIN;
    int64_t M = 1;
    while (M < 2 * n - 1) M <<= 1;
    p->M = M;
    p->inner_f = rfp_plan_new(M, 0);
    p->inner_i = rfp_plan_new(M, 1);
    p->chirp = (cpx*)malloc(sizeof(cpx) * (size_t)n);
    for (int64_t k = 0; k < n; ++k) {
        const unsigned __int128 r = ((unsigned __int128)k * (unsigned __int128)k) % (unsigned __int128)(2 * n);
        p->chirp[k] = twiddle((int64_t)r, 2 * n, inverse);
    }
    p->bspec = (cpx*)calloc((size_t)M, sizeof(cpx));
    for (int64_t k = 0; k < n; ++k) {
        cpx c = p->chirp[k];
        c.im = -c.im;
        p->bspec[k] = c;
        if (k > 0) p->bspec[M - k] = c;
    }
    extern void rfp_plan_process(const rfp_plan*, cpx*);
    rfp_plan_process(p->inner_f, p->bspec);
    for (int64_t k = 0; k < M; ++k) { p->bspec[k].re /= (double)M; p->bspec[k].im /= (double)M; }
    return p;
}

void rfp_plan_free(rfp_plan* p) {
    if (!p) return;
    free(p->tw); free(p->rev); free(p->tw_pm); free(p->bfly); free(p->chirp); free(p->bspec);
    rfp_plan_free(p->sub); rfp_plan_free(p->inner_f); rfp_plan_free(p->inner_i);
    free(p);
}

/* ------------------------------------------------------------------ process */

static void radix4_process(const rfp_plan* p, cpx* buf) {
    const int64_t n = p->n;
    cpx* tmp = (cpx*)malloc(sizeof(cpx) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) tmp[p->rev[i]] = buf[i];
    memcpy(buf, tmp, sizeof(cpx) * (size_t)n);
    free(tmp);
    int lg = 0;
    while (((int64_t)1 << lg) < n) ++lg;
    int64_t len = 1;
    if (lg & 1) { /* radix-2 level */
        for (int64_t i = 0; i < n; i += 2) {
            cpx a = buf[i], b = buf[i + 1];
            buf[i] = cadd(a, b);
            buf[i + 1] = csub(a, b);
        }
        len = 2;
    }
    const double s = p->inverse ? 1.0 : -1.0; /* multiply by -i (fwd) or +i (inv) */
    while (len < n) {
        const int64_t span = len * 4, stride = n / span;
        for (int64_t base = 0; base < n; base += span) {
            for (int64_t k = 0; k < len; ++k) {
                cpx a0 = buf[base + k];
                cpx a1 = cmul(buf[base + k + len], p->tw[k * stride]);
                cpx a2 = cmul(buf[base + k + 2 * len], p->tw[2 * k * stride]);
                cpx a3 = cmul(buf[base + k + 3 * len], p->tw[3 * k * stride]);
                cpx t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), d = csub(a1, a3);
                /* t3 = d * (s*i): forward s = -1 -> -i*d = (d.im, -d.re) */
                cpx t3 = {-s * d.im, s * d.re};
                buf[base + k] = cadd(t0, t2);
                buf[base + k + len] = cadd(t1, t3);
                buf[base + k + 2 * len] = csub(t0, t2);
                buf[base + k + 3 * len] = csub(t1, t3);
            }
        }
        len = span;
    }
}

void rfp_plan_process(const rfp_plan* p, cpx* buf);

static void mixed_process(const rfp_plan* p, cpx* buf) {
    /* n = P*m: x[j + P*i] -> P sub-transforms of length m (decimation in time), twiddle, P-point butterflies */
    const int64_t P = p->p, m = p->m, n = p->n;
    cpx* tmp = (cpx*)malloc(sizeof(cpx) * (size_t)n);
    for (int64_t j = 0; j < P; ++j)
        for (int64_t i = 0; i < m; ++i) tmp[j * m + i] = buf[j + P * i];
    for (int64_t j = 0; j < P; ++j) rfp_plan_process(p->sub, tmp + j * m);
    cpx col[32];
    for (int64_t k = 0; k < m; ++k) {
        for (int64_t j = 0; j < P; ++j) col[j] = cmul(tmp[j * m + k], p->tw_pm[j * m + k]);
        for (int64_t a = 0; a < P; ++a) {
            cpx acc = {0.0, 0.0};
            for (int64_t b = 0; b < P; ++b) acc = cadd(acc, cmul(col[b], p->bfly[a * P + b]));
            buf[k + a * m] = acc;
        }
    }
    free(tmp);
}

static void bluestein_process(const rfp_plan* p, cpx* buf) {
    const int64_t n = p->n, M = p->M;
    cpx* a = (cpx*)calloc((size_t)M, sizeof(cpx));
    for (int64_t k = 0; k < n; ++k) a[k] = cmul(buf[k], p->chirp[k]);
    rfp_plan_process(p->inner_f, a);
    for (int64_t k = 0; k < M; ++k) a[k] = cmul(a[k], p->bspec[k]);
    rfp_plan_process(p->inner_i, a);
    for (int64_t k = 0; k < n; ++k) buf[k] = cmul(a[k], p->chirp[k]);
    free(a);
}

void rfp_plan_process(const rfp_plan* p, cpx* buf) {
    switch (p->alg) {
        case ALG_DFT1: return;
        case ALG_RADIX4: radix4_process(p, buf); return;
        case ALG_MIXED: mixed_process(p, buf); return;
        default: bluestein_process(p, buf); return;
    }
}

/* `rows` contiguous transforms of length n, one plan (like fftn's per-axis plan) */
int rfp_process(void* data, int64_t rows, int64_t n, int inverse) {
    if (rows <= 0 || n <= 0) return -1;
    rfp_plan* p = rfp_plan_new(n, inverse);
    cpx* b = (cpx*)data;
    for (int64_t r = 0; r < rows; ++r) rfp_plan_process(p, b + r * n);
    rfp_plan_free(p);
    return 0;
}

/* ------------------------------------------------- reference call patterns (timed baseline) */

/* fft(x, Some(n)) on complex input: to_complex Vec, resize, FftPlanner::new + plan, copy,
 * process, copy back (fft/algorithms.rs:131-176) */
int rfp_ref_fft(const double* x, int64_t len, int64_t n, int inverse, double* out) {
    cpx* data = (cpx*)calloc((size_t)n, sizeof(cpx));                 /* alloc #1 */
    memcpy(data, x, sizeof(cpx) * (size_t)(len < n ? len : n));
    rfp_plan* p = rfp_plan_new(n, inverse);                           /* re-planned every call */
    cpx* buffer = (cpx*)malloc(sizeof(cpx) * (size_t)n);              /* alloc #2 */
    memcpy(buffer, data, sizeof(cpx) * (size_t)n);
    rfp_plan_process(p, buffer);
    const double sc = inverse ? 1.0 / (double)n : 1.0;
    for (int64_t i = 0; i < n; ++i) { out[2 * i] = buffer[i].re * sc; out[2 * i + 1] = buffer[i].im * sc; } /* alloc #3 */
    rfp_plan_free(p);
    free(buffer);
    free(data);
    return 0;
}

typedef struct {
    const double* x; double* out; int64_t r0, r1, n; int inverse_real; /* 0 rfft, 1 irfft, 2 fft, 3 ifft */
} row_job;

static void rfft_one_row(const double* row, int64_t n, double* out) {
    /* rfft(&row, None): to_complex, fft(n), keep n/2+1 (rfft.rs:39-59) */
    cpx* data = (cpx*)malloc(sizeof(cpx) * (size_t)n);
    for (int64_t i = 0; i < n; ++i) { data[i].re = row[i]; data[i].im = 0.0; }
    rfp_plan* p = rfp_plan_new(n, 0);
    cpx* buffer = (cpx*)malloc(sizeof(cpx) * (size_t)n);
    memcpy(buffer, data, sizeof(cpx) * (size_t)n);
    rfp_plan_process(p, buffer);
    memcpy(out, buffer, sizeof(cpx) * (size_t)(n / 2 + 1));
    rfp_plan_free(p);
    free(buffer);
    free(data);
}

static void irfft_one_row(const double* spec, int64_t n, double* out) {
    /* irfft(&spec, Some(n)): Hermitian extension, ifft(n) with 1/n, real part (rfft.rs:92-178) */
    const int64_t h = n / 2 + 1;
    const cpx* s = (const cpx*)spec;
    cpx* full = (cpx*)malloc(sizeof(cpx) * (size_t)n);
    memcpy(full, s, sizeof(cpx) * (size_t)h);
    int64_t w = h;
    const int64_t start = (n % 2 == 0) ? h - 1 : h;
    for (int64_t i = start - 1; i >= 1 && w < n; --i) { full[w].re = s[i].re; full[w].im = -s[i].im; ++w; }
    for (; w < n; ++w) { full[w].re = 0; full[w].im = 0; }
    rfp_plan* p = rfp_plan_new(n, 1);
    cpx* buffer = (cpx*)malloc(sizeof(cpx) * (size_t)n);
    memcpy(buffer, full, sizeof(cpx) * (size_t)n);
    rfp_plan_process(p, buffer);
    for (int64_t i = 0; i < n; ++i) out[i] = buffer[i].re / (double)n;
    rfp_plan_free(p);
    free(buffer);
    free(full);
}

static void* row_worker(void* arg) {
    row_job* j = (row_job*)arg;
    const int64_t n = j->n, h = n / 2 + 1;
    for (int64_t r = j->r0; r < j->r1; ++r) {
        if (j->inverse_real >= 2)
            rfp_ref_fft(j->x + 2 * r * n, n, n, j->inverse_real == 3, j->out + 2 * r * n);
        else if (j->inverse_real)
            irfft_one_row(j->x + 2 * r * h, n, j->out + r * n);
        else
            rfft_one_row(j->x + r * n, n, j->out + 2 * r * h);
    }
    return NULL;
}

static int rows_threaded(const double* x, int64_t rows, int64_t n, double* out, int nthreads, int inverse_real) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    row_job jobs[256];
    for (int t = 0; t < nthreads; ++t) {
        jobs[t].x = x; jobs[t].out = out; jobs[t].n = n; jobs[t].inverse_real = inverse_real;
        jobs[t].r0 = rows * t / nthreads;
        jobs[t].r1 = rows * (t + 1) / nthreads;
        if (nthreads == 1) { row_worker(&jobs[t]); return 0; }
        pthread_create(&th[t], NULL, row_worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    return 0;
}

/* loop of fft(&row, Some(n)) / ifft over the rows of a [rows][n] complex batch */
int rfp_ref_c2c_rows(const double* x, int64_t rows, int64_t n, double* out, int nthreads, int inverse) {
    return rows_threaded(x, rows, n, out, nthreads, inverse ? 3 : 2);
}

int rfp_ref_rfft_rows(const double* x, int64_t rows, int64_t n, double* out, int nthreads) {
    return rows_threaded(x, rows, n, out, nthreads, 0);
}

int rfp_ref_irfft_rows(const double* spec, int64_t rows, int64_t n, double* out, int nthreads) {
    return rows_threaded(spec, rows, n, out, nthreads, 1);
}

/* fftn(&a, None, axes, None): per axis one plan, per lane gather -> process -> scatter
 * (fft/algorithms.rs:667-690); complex input, in place on `data` (C order). */
int rfp_ref_fftn(double* data, int32_t ndim, const int64_t* shape, const int32_t* axes, int32_t naxes, int inverse) {
    int64_t total = 1;
    for (int d = 0; d < ndim; ++d) total *= shape[d];
    cpx* a = (cpx*)data;
    for (int t = 0; t < naxes; ++t) {
        const int ax = axes[t];
        const int64_t n = shape[ax];
        int64_t inner = 1;
        for (int d = ax + 1; d < ndim; ++d) inner *= shape[d];
        const int64_t outer = total / (n * inner);
        rfp_plan* p = rfp_plan_new(n, inverse);
        cpx* buffer = (cpx*)malloc(sizeof(cpx) * (size_t)n);
        for (int64_t o = 0; o < outer; ++o)
            for (int64_t i = 0; i < inner; ++i) {
                cpx* lane = a + o * n * inner + i;
                for (int64_t k = 0; k < n; ++k) buffer[k] = lane[k * inner];
                rfp_plan_process(p, buffer);
                for (int64_t k = 0; k < n; ++k) lane[k * inner] = buffer[k];
            }
        free(buffer);
        rfp_plan_free(p);
    }
    return 0;
}
