// scirs2_fft_cuda.hpp — header-only C++ mirror of the scirs2-fft interface for the hot path, on top
// of the C ABI (scirs2_fft_cuda.h).  Same names, argument meaning and error behaviour as the Rust
// reference (scirs2-fft/src/fft/algorithms.rs, rfft.rs, backend.rs, plan_cache.rs); `Option<T>` is
// std::optional<T>, `FFTResult<T>` is "T or throw FFTError".  No arithmetic happens here.
#pragma once
#include <complex>
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "scirs2_fft_cuda.h"

namespace scirs2_fft_cuda {

using Complex64 = std::complex<double>;

// FFTError, scirs2-fft/src/error.rs:7-46
struct FFTError : std::runtime_error {
    enum Kind { Computation, Dimension, Value, NotImplemented, IO, Backend, Plan, Communication, Memory } kind;
    FFTError(Kind k, const std::string& m) : std::runtime_error(m), kind(k) {}
};

inline void check(int rc) {
    if (rc >= 0) return;
    static const FFTError::Kind kinds[] = {FFTError::Computation, FFTError::Dimension,      FFTError::Value,
                                           FFTError::NotImplemented, FFTError::IO,          FFTError::Backend,
                                           FFTError::Plan,        FFTError::Communication,  FFTError::Memory};
    const int i = -rc - 1;
    throw FFTError(i >= 0 && i < 9 ? kinds[i] : FFTError::Computation, sfc_last_error());
}

namespace detail {
template <typename T> struct dtype_of;
template <> struct dtype_of<float> { static constexpr int v = SFC_F32; };
template <> struct dtype_of<double> { static constexpr int v = SFC_F64; };
template <> struct dtype_of<std::complex<float>> { static constexpr int v = SFC_C64; };
template <> struct dtype_of<std::complex<double>> { static constexpr int v = SFC_C128; };
inline int64_t next_pow2(int64_t n) { int64_t p = 1; while (p < n) p <<= 1; return p; }
inline int64_t prod(const std::vector<int64_t>& v) { int64_t p = 1; for (auto x : v) p *= x; return p; }
}  // namespace detail

// fft / ifft — fft/algorithms.rs:131-176, 210-263
template <typename T>
std::vector<Complex64> fft(const std::vector<T>& x, std::optional<size_t> n = std::nullopt) {
    const int64_t cap = n ? (int64_t)*n : detail::next_pow2((int64_t)x.size());
    std::vector<Complex64> out((size_t)std::max<int64_t>(cap, 1));
    int64_t len = 0;
    check(sfc_fft(x.data(), (int64_t)x.size(), detail::dtype_of<T>::v, n ? (int64_t)*n : -1,
                  reinterpret_cast<double*>(out.data()), (int64_t)out.size(), &len));
    out.resize((size_t)len);
    return out;
}
template <typename T>
std::vector<Complex64> ifft(const std::vector<T>& x, std::optional<size_t> n = std::nullopt) {
    const int64_t cap = n ? (int64_t)*n : detail::next_pow2((int64_t)x.size());
    std::vector<Complex64> out((size_t)std::max<int64_t>(cap, 1));
    int64_t len = 0;
    check(sfc_ifft(x.data(), (int64_t)x.size(), detail::dtype_of<T>::v, n ? (int64_t)*n : -1,
                   reinterpret_cast<double*>(out.data()), (int64_t)out.size(), &len));
    out.resize((size_t)len);
    return out;
}
// rfft / irfft — rfft.rs:39-59, 92-178
template <typename T>
std::vector<Complex64> rfft(const std::vector<T>& x, std::optional<size_t> n = std::nullopt) {
    const int64_t nv = n ? (int64_t)*n : (int64_t)x.size();
    std::vector<Complex64> out((size_t)(nv / 2 + 1));
    int64_t len = 0;
    check(sfc_rfft(x.data(), (int64_t)x.size(), detail::dtype_of<T>::v, n ? (int64_t)*n : -1,
                   reinterpret_cast<double*>(out.data()), (int64_t)out.size(), &len));
    out.resize((size_t)len);
    return out;
}
template <typename T>
std::vector<double> irfft(const std::vector<T>& x, std::optional<size_t> n = std::nullopt) {
    const int64_t nv = n ? (int64_t)*n : 2 * ((int64_t)x.size() - 1);
    std::vector<double> out((size_t)std::max<int64_t>(nv, 1));
    int64_t len = 0;
    check(sfc_irfft(x.data(), (int64_t)x.size(), detail::dtype_of<T>::v, n ? (int64_t)*n : -1, out.data(),
                    (int64_t)out.size(), &len));
    out.resize((size_t)len);
    return out;
}

// N-D arrays: C-order data + shape (the ndarray `ArrayD` of the reference)
template <typename T>
struct ArrayD {
    std::vector<int64_t> shape;
    std::vector<T> data;
};

// fftn / ifftn — fft/algorithms.rs:576-706, 757-890 (overwrite_x and workers are accepted and ignored, :581-582)
template <typename T>
ArrayD<Complex64> fftn(const ArrayD<T>& x, const std::optional<std::vector<int64_t>>& shape = std::nullopt,
                       const std::optional<std::vector<int64_t>>& axes = std::nullopt, const char* norm = nullptr,
                       bool inverse = false) {
    if (shape && shape->size() != x.shape.size())
        throw FFTError(FFTError::Value, "Output shape must have the same number of dimensions as input");
    ArrayD<Complex64> out;
    out.shape = shape ? *shape : x.shape;
    out.data.resize((size_t)std::max<int64_t>(detail::prod(out.shape), 1));
    std::vector<int64_t> oshape(x.shape.size());
    auto fn = inverse ? sfc_ifftn : sfc_fftn;
    check(fn(x.data.data(), (int32_t)x.shape.size(), x.shape.data(), detail::dtype_of<T>::v,
             shape ? shape->data() : nullptr, axes ? axes->data() : nullptr, axes ? (int32_t)axes->size() : 0, norm,
             reinterpret_cast<double*>(out.data.data()), (int64_t)out.data.size(), oshape.data()));
    out.shape = oshape;
    return out;
}
template <typename T>
ArrayD<Complex64> ifftn(const ArrayD<T>& x, const std::optional<std::vector<int64_t>>& shape = std::nullopt,
                        const std::optional<std::vector<int64_t>>& axes = std::nullopt, const char* norm = nullptr) {
    return fftn(x, shape, axes, norm, true);
}
// fft2 / ifft2 — fft/algorithms.rs:293-401, 439-541
template <typename T>
ArrayD<Complex64> fft2(const ArrayD<T>& x, const std::optional<std::pair<size_t, size_t>>& shape = std::nullopt,
                       const std::optional<std::pair<int, int>>& axes = std::nullopt, const char* norm = nullptr,
                       bool inverse = false) {
    if (x.shape.size() != 2) throw FFTError(FFTError::Dimension, "expected a 2-D array");
    int64_t sh[2] = {shape ? (int64_t)shape->first : x.shape[0], shape ? (int64_t)shape->second : x.shape[1]};
    int32_t ax[2] = {axes ? axes->first : 0, axes ? axes->second : 1};
    ArrayD<Complex64> out;
    out.shape = {sh[0], sh[1]};
    out.data.resize((size_t)std::max<int64_t>(sh[0] * sh[1], 1));
    int64_t os[2];
    auto fn = inverse ? sfc_ifft2 : sfc_fft2;
    check(fn(x.data.data(), x.shape[0], x.shape[1], detail::dtype_of<T>::v, shape ? sh : nullptr, axes ? ax : nullptr,
             norm, reinterpret_cast<double*>(out.data.data()), (int64_t)out.data.size(), os));
    return out;
}

// ---- consumers of the path: dct.rs, dst.rs, hartley.rs, hfft/*.rs, lib.rs::hilbert -----------------
enum class DCTType : int { Type1 = 1, Type2 = 2, Type3 = 3, Type4 = 4 };  // dct.rs:13-23
enum class DSTType : int { Type1 = 1, Type2 = 2, Type3 = 3, Type4 = 4 };  // dst.rs:13-22
namespace detail {
inline std::vector<double> trig1d(decltype(&sfc_dct) fn, const std::vector<double>& x, int type, bool inverse, const char* norm) {
    std::vector<double> out(std::max<size_t>(x.size(), 1));
    const int64_t shape[1] = {(int64_t)x.size()};
    const int32_t axes[1] = {0};
    check(fn(x.data(), 1, shape, axes, 1, type, inverse ? 1 : 0, norm, out.data()));
    out.resize(x.size());
    return out;
}
inline ArrayD<double> trignd(decltype(&sfc_dct) fn, const ArrayD<double>& x, int type, bool inverse, const char* norm,
                             const std::optional<std::vector<int32_t>>& axes) {
    ArrayD<double> out{x.shape, std::vector<double>(x.data.size())};
    check(fn(x.data.data(), (int32_t)x.shape.size(), x.shape.data(), axes ? axes->data() : nullptr,
             axes ? (int32_t)axes->size() : 0, type, inverse ? 1 : 0, norm, out.data.data()));
    return out;
}
}  // namespace detail
// dct / idct / dctn / idctn — dct.rs:56-420 (norm: "ortho" or nullptr)
inline std::vector<double> dct(const std::vector<double>& x, std::optional<DCTType> t = std::nullopt, const char* norm = nullptr) {
    return detail::trig1d(sfc_dct, x, (int)t.value_or(DCTType::Type2), false, norm);
}
inline std::vector<double> idct(const std::vector<double>& x, std::optional<DCTType> t = std::nullopt, const char* norm = nullptr) {
    return detail::trig1d(sfc_dct, x, (int)t.value_or(DCTType::Type2), true, norm);
}
inline ArrayD<double> dctn(const ArrayD<double>& x, std::optional<DCTType> t = std::nullopt, const char* norm = nullptr,
                           const std::optional<std::vector<int32_t>>& axes = std::nullopt) {
    return detail::trignd(sfc_dct, x, (int)t.value_or(DCTType::Type2), false, norm, axes);
}
inline ArrayD<double> idctn(const ArrayD<double>& x, std::optional<DCTType> t = std::nullopt, const char* norm = nullptr,
                            const std::optional<std::vector<int32_t>>& axes = std::nullopt) {
    return detail::trignd(sfc_dct, x, (int)t.value_or(DCTType::Type2), true, norm, axes);
}
// dst / idst / dstn / idstn — dst.rs:48-405
inline std::vector<double> dst(const std::vector<double>& x, std::optional<DSTType> t = std::nullopt, const char* norm = nullptr) {
    return detail::trig1d(sfc_dst, x, (int)t.value_or(DSTType::Type2), false, norm);
}
inline std::vector<double> idst(const std::vector<double>& x, std::optional<DSTType> t = std::nullopt, const char* norm = nullptr) {
    return detail::trig1d(sfc_dst, x, (int)t.value_or(DSTType::Type2), true, norm);
}
inline ArrayD<double> dstn(const ArrayD<double>& x, std::optional<DSTType> t = std::nullopt, const char* norm = nullptr,
                           const std::optional<std::vector<int32_t>>& axes = std::nullopt) {
    return detail::trignd(sfc_dst, x, (int)t.value_or(DSTType::Type2), false, norm, axes);
}
inline ArrayD<double> idstn(const ArrayD<double>& x, std::optional<DSTType> t = std::nullopt, const char* norm = nullptr,
                            const std::optional<std::vector<int32_t>>& axes = std::nullopt) {
    return detail::trignd(sfc_dst, x, (int)t.value_or(DSTType::Type2), true, norm, axes);
}
// dht / idht — hartley.rs:37-130;  hilbert — lib.rs:437-516;  hfft / ihfft — hfft/*.rs
inline std::vector<double> dht(const std::vector<double>& x) {
    std::vector<double> out(std::max<size_t>(x.size(), 1));
    check(sfc_dht(x.data(), (int64_t)x.size(), out.data()));
    out.resize(x.size());
    return out;
}
inline std::vector<double> idht(const std::vector<double>& h) {
    std::vector<double> out(std::max<size_t>(h.size(), 1));
    check(sfc_idht(h.data(), (int64_t)h.size(), out.data()));
    out.resize(h.size());
    return out;
}
inline std::vector<Complex64> hilbert(const std::vector<double>& x) {
    std::vector<Complex64> out(std::max<size_t>(x.size(), 1));
    check(sfc_hilbert(x.data(), (int64_t)x.size(), reinterpret_cast<double*>(out.data())));
    out.resize(x.size());
    return out;
}
template <typename T>
std::vector<double> hfft(const std::vector<T>& x, std::optional<size_t> n = std::nullopt) {
    std::vector<double> out((size_t)std::max<int64_t>(n ? (int64_t)*n : (int64_t)x.size(), 1));
    int64_t len = 0;
    check(sfc_hfft(x.data(), (int64_t)x.size(), detail::dtype_of<T>::v, n ? (int64_t)*n : -1, out.data(), (int64_t)out.size(), &len));
    out.resize((size_t)len);
    return out;
}
inline std::vector<Complex64> ihfft(const std::vector<double>& x, std::optional<size_t> n = std::nullopt) {
    std::vector<Complex64> out((size_t)std::max<int64_t>(n ? (int64_t)*n : (int64_t)x.size(), 1));
    int64_t len = 0;
    check(sfc_ihfft(x.data(), (int64_t)x.size(), n ? (int64_t)*n : -1, reinterpret_cast<double*>(out.data()), (int64_t)out.size(), &len));
    out.resize((size_t)len);
    return out;
}

// czt — czt.rs:279-303 (1-D; w / a as in CZT::new, czt.rs:64-131)
inline std::vector<Complex64> czt(const std::vector<Complex64>& x, std::optional<size_t> m = std::nullopt,
                                  std::optional<Complex64> w = std::nullopt, std::optional<Complex64> a = std::nullopt) {
    const int64_t mm = m ? (int64_t)*m : (int64_t)x.size();
    std::vector<Complex64> out((size_t)std::max<int64_t>(mm, 1));
    const Complex64 wv = w.value_or(Complex64(0.0, 0.0)), av = a.value_or(Complex64(1.0, 0.0));
    check(sfc_czt(reinterpret_cast<const double*>(x.data()), 1, (int64_t)x.size(), mm, w ? 1 : 0, wv.real(), wv.imag(), av.real(),
                  av.imag(), reinterpret_cast<double*>(out.data())));
    out.resize((size_t)mm);
    return out;
}

// scirs2-signal spectral.rs:257-410 (welch) as one framed, batched device transform.  `window` holds the
// nperseg samples the caller's get_window produced; detrend: "none" | "constant" | "linear".  Returns the averaged density
// for the first nfft/2 + nfft%2 bins (multiply by fs for the "spectrum" scaling, spectral.rs:402-407).
inline std::vector<double> welch_psd(const std::vector<double>& x, double fs, const std::vector<double>& window, size_t noverlap,
                                     size_t nfft, const std::string& detrend = "constant") {
    const size_t nperseg = window.size();
    if (x.empty()) throw FFTError(FFTError::Value, "Input array is empty");
    if (nperseg == 0 || nfft < nperseg || noverlap >= nperseg) throw FFTError(FFTError::Value, "bad nperseg / noverlap / nfft");
    const int d = detrend == "none" ? 0 : detrend == "constant" ? 1 : detrend == "linear" ? 2 : -1;
    if (d < 0) throw FFTError(FFTError::Value, "Unknown detrend option: " + detrend);
    const size_t step = nperseg - noverlap;
    const size_t segs = x.size() >= noverlap ? (x.size() - noverlap) / step : 0;
    if (segs < 1) throw FFTError(FFTError::Value, "Not enough data points for given nperseg and noverlap");
    size_t P = 1;
    while (P < nfft) P <<= 1;
    double w2 = 0.0;
    for (double w : window) w2 += w * w;
    std::vector<double> psd(nfft / 2 + nfft % 2);
    check(sfc_signal_spectra(x.data(), (int64_t)x.size(), window.data(), (int64_t)nperseg, (int64_t)step, (int64_t)segs, (int64_t)P, d, 1,
                             (int64_t)psd.size(), 1.0 / w2 / (fs * (double)nperseg), psd.data()));
    for (double& v : psd) v /= (double)segs;
    return psd;
}

// PlanCache — plan_cache.rs:28-235 (the cache itself lives in the library)
struct CacheStats { uint64_t hit_count, miss_count; double hit_rate; uint64_t size, max_size; };
class PlanCache {
   public:
    void set_enabled(bool e) { check(sfc_cache_set_enabled(e ? 1 : 0)); }
    bool is_enabled() const { return sfc_cache_is_enabled() != 0; }
    void clear() { check(sfc_cache_clear()); }
    CacheStats get_stats() const {
        sfc_cache_stats s;
        check(sfc_cache_get_stats(&s));
        return {s.hit_count, s.miss_count, s.hit_rate, s.size, s.max_size};
    }
};
inline PlanCache& get_global_cache() { static PlanCache c; return c; }

// trait FftBackend — backend.rs:14-48
class FftBackend {
   public:
    virtual ~FftBackend() = default;
    virtual const char* name() const = 0;
    virtual const char* description() const = 0;
    virtual bool is_available() const = 0;
    virtual void fft(const std::vector<Complex64>& in, std::vector<Complex64>& out) const = 0;
    virtual void ifft(const std::vector<Complex64>& in, std::vector<Complex64>& out) const = 0;
    virtual void fft_sized(const std::vector<Complex64>& in, std::vector<Complex64>& out, size_t size) const = 0;
    virtual void ifft_sized(const std::vector<Complex64>& in, std::vector<Complex64>& out, size_t size) const = 0;
    virtual bool supports_feature(const std::string& f) const = 0;
};

class CudaFftBackend final : public FftBackend {
   public:
    const char* name() const override { return sfc_backend_name(); }
    const char* description() const override { return sfc_backend_description(); }
    bool is_available() const override { return sfc_is_available() != 0; }
    void fft(const std::vector<Complex64>& in, std::vector<Complex64>& out) const override {
        fft_sized(in, out, in.size());
    }
    void ifft(const std::vector<Complex64>& in, std::vector<Complex64>& out) const override {
        ifft_sized(in, out, in.size());
    }
    void fft_sized(const std::vector<Complex64>& in, std::vector<Complex64>& out, size_t size) const override {
        check(sfc_backend_fft_sized(reinterpret_cast<const double*>(in.data()), (int64_t)in.size(),
                                    reinterpret_cast<double*>(out.data()), (int64_t)out.size(), (int64_t)size));
    }
    void ifft_sized(const std::vector<Complex64>& in, std::vector<Complex64>& out, size_t size) const override {
        check(sfc_backend_ifft_sized(reinterpret_cast<const double*>(in.data()), (int64_t)in.size(),
                                     reinterpret_cast<double*>(out.data()), (int64_t)out.size(), (int64_t)size));
    }
    bool supports_feature(const std::string& f) const override { return sfc_backend_supports_feature(f.c_str()) != 0; }
};

// BackendManager — backend.rs:163-281
class BackendManager {
   public:
    BackendManager() { backends_["cuda_fft"] = std::make_shared<CudaFftBackend>(); }
    void register_backend(const std::string& name, std::shared_ptr<FftBackend> b) {
        std::lock_guard<std::mutex> lk(mu_);
        if (backends_.count(name)) throw FFTError(FFTError::Value, "Backend '" + name + "' already exists");
        backends_[name] = std::move(b);
    }
    void set_backend(const std::string& name) {
        std::lock_guard<std::mutex> lk(mu_);
        auto it = backends_.find(name);
        if (it == backends_.end()) throw FFTError(FFTError::Value, "Backend '" + name + "' not found");
        if (!it->second->is_available()) throw FFTError(FFTError::Value, "Backend '" + name + "' is not available");
        current_ = name;
    }
    std::shared_ptr<FftBackend> get_backend() {
        std::lock_guard<std::mutex> lk(mu_);
        return backends_.at(current_);
    }
    std::string get_backend_name() {
        std::lock_guard<std::mutex> lk(mu_);
        return current_;
    }

   private:
    std::mutex mu_;
    std::map<std::string, std::shared_ptr<FftBackend>> backends_;
    std::string current_ = "cuda_fft";
};
inline BackendManager& get_backend_manager() { static BackendManager m; return m; }

// ---------------------------------------------------------------- distributed.rs:18-103, 115-362
// `trait Communicator` for the GPUs of one node and the slab-decomposed 3-D transform, both inside the library.
enum class DecompositionStrategy { Replicated = SFC_DECOMP_REPLICATED, BatchSplit = SFC_DECOMP_BATCH_SPLIT, Slab = SFC_DECOMP_SLAB };
enum class SlabLayout { Transposed = SFC_SLAB_TRANSPOSED, Natural = SFC_SLAB_NATURAL };

class Communicator {
   public:
    // one process per GPU: every rank passes the same job-unique name
    static Communicator rank(const std::string& name, int rank, int world, int device) {
        sfc_comm* h = nullptr;
        check(sfc_comm_init_rank(&h, name.c_str(), rank, world, device));
        return Communicator(h);
    }
    // one process driving ngpu GPUs (0 = all visible)
    static Communicator local(int ngpu = 0) {
        sfc_comm* h = nullptr;
        check(sfc_comm_init_local(&h, ngpu, nullptr));
        return Communicator(h);
    }
    Communicator(Communicator&& o) noexcept : h_(o.h_) { o.h_ = nullptr; }
    Communicator(const Communicator&) = delete;
    ~Communicator() { if (h_) sfc_comm_destroy(h_); }
    int size() const { return sfc_comm_size(h_); }   // Communicator::size, distributed.rs:99
    int rank() const { return sfc_comm_rank(h_); }   // Communicator::rank, :102
    void barrier() const { check(sfc_comm_barrier(h_)); }  // Communicator::barrier, :96
    sfc_comm* handle() const { return h_; }

   private:
    explicit Communicator(sfc_comm* h) : h_(h) {}
    sfc_comm* h_;
};

// DistributedFFT::distributed_fft (distributed.rs:115-163) for 3-D complex volumes
class DistributedFFT {
   public:
    DistributedFFT(const Communicator& comm, const std::vector<int64_t>& shape, bool inverse = false,
                   SlabLayout layout = SlabLayout::Natural, double scale = 1.0) {
        if (shape.size() != 3) throw FFTError(FFTError::Value, "slab decomposition: 3-D complex transforms only");
        sfc_dist_desc d{};
        d.base.ndim = 3;
        d.base.naxes = 3;
        for (int i = 0; i < 3; ++i) { d.base.shape[i] = shape[(size_t)i]; d.base.axes[i] = i; }
        d.base.kind = SFC_C2C;
        d.base.prec = SFC_PREC_F64;
        d.base.direction = inverse ? SFC_INVERSE : SFC_FORWARD;
        d.base.scale = scale;
        d.decomposition = SFC_DECOMP_SLAB;
        d.layout = (int)layout;
        check(sfc_dist_plan_create(&h_, comm.handle(), &d));
        check(sfc_dist_plan_get_info(h_, &info));
    }
    DistributedFFT(const DistributedFFT&) = delete;
    ~DistributedFFT() { if (h_) sfc_dist_plan_destroy(h_); }
    // host buffers: local mode = the whole C-order volume, rank mode = this rank's slab in and its share out
    void execute(const Complex64* in, Complex64* out) const { check(sfc_dist_exec_host(h_, in, out)); }
    sfc_dist_info info{};

   private:
    sfc_dist_plan* h_ = nullptr;
};

// fftn / ifftn / execute_batch of THIS process run over ngpu GPUs from now on (-1: all visible, 1: back to one)
inline void set_num_gpus(int ngpu) { check(sfc_set_num_gpus(ngpu)); }

}  // namespace scirs2_fft_cuda
