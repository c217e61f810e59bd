// Exercises include/scirs2_fft_cuda.hpp the way the reference's tests exercise scirs2-fft
// (doctests of fft/algorithms.rs, rfft.rs:926-966, backend.rs:350-386).  Exit code 0 = pass.
// Without a CUDA device it checks that every call fails with FFTError::Backend (no CPU fallback).
#include <cmath>
#include <cstdio>
#include "scirs2_fft_cuda.hpp"
using namespace scirs2_fft_cuda;
#define REQUIRE(c) do { if (!(c)) { std::printf("FAILED: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)
int main() {
    std::vector<double> sig{1.0, 2.0, 3.0, 4.0};
    if (!sfc_is_available()) {
        try { fft(sig); } catch (const FFTError& e) { REQUIRE(e.kind == FFTError::Backend); std::puts("no device: BackendError ok"); return 0; }
        return 1;
    }
    auto s = fft(sig);
    REQUIRE(std::abs(s[0].real() - 10.0) < 1e-10 && std::abs(s[0].imag()) < 1e-10);
    auto r = ifft(s);
    for (int i = 0; i < 4; ++i) REQUIRE(std::abs(r[i].real() - sig[i]) < 1e-10 && std::abs(r[i].imag()) < 1e-10);
    auto sp = rfft(sig);
    REQUIRE(sp.size() == 3 && std::abs(sp[0].real() - 10.0) < 1e-10);
    auto back = irfft(sp, 4);
    for (int i = 0; i < 4; ++i) REQUIRE(std::abs(back[i] - sig[i]) < 1e-10);
    REQUIRE(fft(std::vector<double>(100, 1.0)).size() == 128);  // next-power-of-two padding quirk
    ArrayD<double> a{{2, 2}, {1, 2, 3, 4}};
    REQUIRE(std::abs(fft2(a).data[0].real() - 10.0) < 1e-10);
    ArrayD<double> v{{2, 2, 2}, {0, 1, 2, 3, 4, 5, 6, 7}};
    auto rt = ifftn(fftn(v));
    for (int i = 0; i < 8; ++i) REQUIRE(std::abs(rt.data[i].real() - v.data[i]) < 1e-10);
    try { fftn(v, std::nullopt, std::vector<int64_t>{3}); return 1; } catch (const FFTError& e) {
        REQUIRE(e.kind == FFTError::Value && std::string(e.what()) == "Axis 3 out of bounds for array of dimension 3");
    }
    auto b = get_backend_manager().get_backend();
    REQUIRE(std::string(b->name()) == "cuda_fft" && b->supports_feature("gpu_acceleration"));
    std::vector<Complex64> imp(8, 0.0), out(8);
    imp[0] = 1.0;
    b->fft(imp, out);
    for (auto& c : out) REQUIRE(std::abs(std::abs(c) - 1.0) < 1e-10);
    try { b->fft_sized(imp, out, 4); return 1; } catch (const FFTError& e) { REQUIRE(e.kind == FFTError::Value); }
    // consumers (dct.rs:757-768, 843-862; hartley.rs:216-231)
    auto c = dct(sig, DCTType::Type2, "ortho");
    auto rc = idct(c, DCTType::Type2, "ortho");
    for (int i = 0; i < 4; ++i) REQUIRE(std::abs(rc[i] - sig[i]) < 1e-10);
    auto cc = dct(std::vector<double>(4, 3.0));
    REQUIRE(std::abs(cc[0]) > 1e-10 && std::abs(cc[1]) < 1e-10 && std::abs(cc[3]) < 1e-10);
    auto hh = idht(dht(sig));
    for (int i = 0; i < 4; ++i) REQUIRE(std::abs(hh[i] - sig[i]) < 1e-10);
    REQUIRE(hilbert(sig).size() == 4 && hfft(sig).size() == 4 && ihfft(sig).size() == 4 && dst(sig).size() == 4);
    try { dct(std::vector<double>{1.0}, DCTType::Type1); return 1; } catch (const FFTError& e) { REQUIRE(e.kind == FFTError::Value); }
    {  // czt with default parameters equals fft (czt.rs:396-410)
        std::vector<Complex64> xc(8);
        for (int i = 0; i < 8; ++i) xc[i] = Complex64((double)i, 0.0);
        auto zc = czt(xc);
        auto fc = fft(xc, 8);
        for (int i = 0; i < 8; ++i) REQUIRE(std::abs(zc[i] - fc[i]) < 1e-10);
    }
    {  // welch (scirs2-signal spectral.rs:257-410): a tone on bin 4 of 64, boxcar, no detrend, no overlap:
       // every segment has |X[4]|^2 = 32^2, so the density is 1024 / sum(w^2) / (fs * nperseg) = 0.25 there and 0 elsewhere
        std::vector<double> tone(640), box(64, 1.0);
        for (int i = 0; i < 640; ++i) tone[i] = std::cos(2.0 * 3.14159265358979323846 * 4.0 * i / 64.0);
        auto psd = welch_psd(tone, 1.0, box, 0, 64, "none");
        REQUIRE(psd.size() == 32);
        for (int k = 0; k < 32; ++k) REQUIRE(std::abs(psd[k] - (k == 4 ? 0.25 : 0.0)) < 1e-12);
        try { welch_psd(tone, 1.0, box, 0, 64, "bogus"); return 1; } catch (const FFTError& e) { REQUIRE(e.kind == FFTError::Value); }
    }
    {  // distributed.rs mirror on however many GPUs are visible (one: the degenerate slab plan)
        auto comm = Communicator::local(1);
        REQUIRE(comm.size() == 1 && comm.rank() == 0);
        DistributedFFT dplan(comm, {8, 16, 4});
        std::vector<Complex64> vin(8 * 16 * 4, Complex64(0.0, 0.0)), vout(vin.size());
        vin[0] = 1.0;  // impulse -> all ones
        dplan.execute(vin.data(), vout.data());
        for (auto& c : vout) REQUIRE(std::abs(c - Complex64(1.0, 0.0)) < 1e-12);
        REQUIRE(dplan.info.world == 1 && dplan.info.num_exchanges == 0);
    }
    std::puts("cpp mirror ok");
    return 0;
}
