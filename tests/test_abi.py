"""CPU: the C-ABI library loads, exports every symbol the header declares, and fails loudly
(no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "scirs2_fft_cuda.h")


def _has_gpu():
    try:
        from scirs_b200 import _lib

        return _lib.load().sfc_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="module")
def lib(build_artifacts):
    from scirs_b200 import _lib

    return _lib.load()


def header_symbols():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sfc_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_are_exported(lib):
    syms = header_symbols()
    assert len(syms) >= 35
    out = subprocess.run(["nm", "-D", "--defined-only", lib._name], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\b(sfc_[a-z0-9_]+)\b", out))
    missing = [s for s in syms if s not in exported]
    assert not missing, f"header declares symbols the library does not export: {missing}"
    # nothing but the C ABI leaks out of the shared object
    leaked = [l for l in out.splitlines() if " T " in l and "sfc_" not in l]
    assert not leaked, leaked


def test_python_binding_covers_header(lib):
    from scirs_b200 import _lib

    assert sorted(_lib.SIGNATURES) == header_symbols()
    assert lib.sfc_abi_version() == 2


def test_rust_ffi_is_generated_from_the_header():
    """rust/scirs2-fft-cuda/src/ffi.rs cannot be compiled here (no rustc): it is at least a mechanical image of the header."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("gen_rust_ffi", os.path.join(ROOT, "tools", "gen_rust_ffi.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    text = gen.generate()
    assert open(gen.OUT).read() == text, "ffi.rs is stale: run python tools/gen_rust_ffi.py"
    rs_fns = set(re.findall(r"pub fn (sfc_[a-z0-9_]+)\(", text))
    assert sorted(rs_fns) == header_symbols()
    # every struct the header defines, field for field
    hdr = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name in re.findall(r"typedef\s+struct\s+(\w+)\s*\{", hdr):
        assert f"pub struct {name} {{" in text


def test_struct_layouts_match_header(lib):
    from scirs_b200 import _lib

    # sfc_desc: int32 ndim (+pad) | int64 shape[8] | int32 naxes, axes[8], kind, prec, direction, flags (+pad) | double | int64[8]
    #           | int32 scatter_parts, reserved | int64 scatter_pitch | int64 axis_in_len, axis_out_len | void* aux_in, aux_out | double scale_dc
    assert C.sizeof(_lib.sfc_desc) == 8 + 64 + 4 * 13 + 4 + 8 + 64 + 8 + 8 + 16 + 16 + 8
    assert C.sizeof(_lib.sfc_plan_info) == 8 * 5 + 8 + 4 * 2
    assert C.sizeof(_lib.sfc_cache_stats) == 40
    assert C.sizeof(_lib.sfc_dist_desc) == C.sizeof(_lib.sfc_desc) + 16
    assert C.sizeof(_lib.sfc_dist_info) == 4 * 6 + 8 * 2 + 64 * 2 + 8 + 4 * 2 + 8 + 8
    # the C compiler's view of the same structs
    import tempfile

    src = '#include "scirs2_fft_cuda.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu\\n", sizeof(sfc_desc), sizeof(sfc_plan_info), sizeof(sfc_dist_desc), sizeof(sfc_dist_info));return 0;}\n'
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "sz.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.dirname(HEADER), os.path.join(td, "sz.c"), "-o", os.path.join(td, "sz")], check=True)
        out = subprocess.run([os.path.join(td, "sz")], capture_output=True, text=True, check=True).stdout.split()
    assert [int(v) for v in out] == [C.sizeof(_lib.sfc_desc), C.sizeof(_lib.sfc_plan_info), C.sizeof(_lib.sfc_dist_desc), C.sizeof(_lib.sfc_dist_info)]


def test_sm100a_only(lib):
    out = subprocess.run(["cuobjdump", "-lelf", lib._name], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_uses_bulk_async_and_fp64_pipe_not_tensor_cores(lib):
    # FFT is bandwidth-bound: no tcgen05/HMMA in the product (north_star), DFMA present
    out = subprocess.run(["cuobjdump", "-sass", lib._name], capture_output=True, text=True).stdout
    assert "DFMA" in out
    assert not re.search(r"\b(UTC\w*MMA|HMMA|DMMA)\b", out)


@pytest.mark.skipif(_has_gpu(), reason="CPU-only behaviour")
def test_no_cpu_fallback(lib):
    import scirs_b200 as sb

    assert lib.sfc_device_count() == 0 and lib.sfc_is_available() == 0
    assert lib.sfc_init(0) == -6
    for call in (lambda: sb.fft(np.ones(8)), lambda: sb.rfft(np.ones(8)), lambda: sb.fft2(np.ones((4, 4))),
                 lambda: sb.fftn(np.ones((2, 2, 2))), lambda: sb.irfftn(np.ones((2, 2, 2)) + 0j),
                 lambda: sb.FftPlan([8], [0]), lambda: sb.rfft_batch(np.ones((2, 64))),
                 # the consumers of the path have no CPU fallback either
                 lambda: sb.dct(np.ones(8)), lambda: sb.idst(np.ones(8), sb.DSTType.Type3, "ortho"), lambda: sb.dctn(np.ones((4, 4))),
                 lambda: sb.dht(np.ones(8)), lambda: sb.dht2(np.ones((4, 4))), lambda: sb.hfft(np.ones(8) + 0j),
                 lambda: sb.ihfft(np.ones(8)), lambda: sb.hilbert(np.ones(8)), lambda: sb.stft(np.ones(64), "hann", 16),
                 lambda: sb.spectrogram(np.ones(64), nperseg=16), lambda: sb.fft2_efficient(np.ones((4, 4))),
                 lambda: sb.fft_streaming(np.ones(8)), lambda: sb.fftn_optimized(np.ones((4, 4))),
                 lambda: sb.fft_inplace(np.ones(8, dtype=np.complex128), np.ones(8, dtype=np.complex128)),
                 lambda: sb.czt(np.ones(8) + 0j),
                 # ... nor do the scirs2-signal callers (8f rank 4)
                 lambda: sb.signal.periodogram(np.arange(16.0)), lambda: sb.signal.welch(np.arange(64.0), nperseg=16),
                 lambda: sb.signal.stft(np.arange(64.0), nperseg=16), lambda: sb.signal.wiener_filter(np.arange(32.0)),
                 lambda: sb.signal.spectral_subtraction(np.arange(128.0)), lambda: sb.signal.psd_wiener_filter(np.arange(32.0)),
                 lambda: sb.signal.StreamingStft(sb.signal.StreamingStftConfig(16, 8)).process_frame(np.ones(32)),
                 lambda: sb.signal.bispectrum(np.arange(64.0), 16), lambda: sb.signal.hilbert(np.arange(16.0)),
                 lambda: sb.signal.wigner_ville(np.arange(16.0)), lambda: sb.signal.constant_q_transform(np.arange(64.0), sb.signal.CqtConfig(f_min=200.0, f_max=400.0, fs=2000.0)),
                 lambda: sb.signal.bicoherence(np.arange(64.0), 16)):
        with pytest.raises(sb.BackendError) as e:
            call()
        assert "no CPU fallback" in str(e.value)
    b = sb.CudaFftBackend()
    assert b.name() == "cuda_fft" and not b.is_available()
    assert b.supports_feature("1d_fft") and b.supports_feature("gpu_acceleration") and not b.supports_feature("x")


def test_cache_controls_work_without_gpu(lib):
    import scirs_b200 as sb

    c = sb.get_global_cache()
    c.configure(128, 3600.0)
    s = c.get_stats()
    assert (s.hit_count, s.miss_count, s.size, s.max_size, s.hit_rate) == (0, 0, 0, 128, 0.0)
    c.set_enabled(False)
    assert not c.is_enabled()
    c.set_enabled(True)
    assert c.is_enabled()
    c.clear()


def test_argument_validation_precedes_device_use(lib):
    import scirs_b200 as sb

    m = sb.get_backend_manager()
    assert m.get_backend_name() == "cuda_fft" and "cuda_fft" in m.list_backends()
    with pytest.raises(sb.ValueError_):
        m.register_backend("cuda_fft", sb.CudaFftBackend())  # backend.rs:184-194
    with pytest.raises(sb.ValueError_):
        m.set_backend("nope")  # backend.rs:203-224
    # size check of fft_sized comes before any device work (backend.rs:96-100)
    x = np.zeros(8, dtype=np.complex128)
    with pytest.raises(sb.ValueError_) as e:
        sb.CudaFftBackend().fft_sized(x, np.zeros(4, dtype=np.complex128), 8)
    assert str(e.value) == "Input and output sizes must match the specified size"
    with pytest.raises(sb.NotImplementedError_):
        sb.ifft2_simd(np.ones((2, 2)))  # simd_fft.rs:99-111


def test_context_and_worker_pool_mirrors(lib, monkeypatch):
    """worker_pool.rs:229-300, context.rs:216-276 (configuration holders; no device needed)"""
    import scirs_b200 as sb
    from scirs_b200 import context

    monkeypatch.setenv("SCIRS2_FFT_WORKERS", "3")
    assert context.WorkerConfig().num_workers == 3  # worker_pool.rs:32-35
    pool = sb.WorkerPool()
    pool.set_workers(5)
    assert pool.get_workers() == 5 and pool.is_enabled()
    assert pool.execute(lambda: 42) == 42 and pool.execute_with_workers(2, lambda: 7) == 7
    assert pool.get_info().thread_name_prefix == "scirs2-fft-worker"
    c = sb.get_global_cache()
    assert c.is_enabled()
    assert sb.without_cache(lambda: c.is_enabled()) is False and c.is_enabled()
    assert sb.with_workers(4, lambda: 11) == 11
    if lib.sfc_is_available():
        assert sb.with_backend("cuda_fft", lambda: sb.get_backend_manager().get_backend_name()) == "cuda_fft"
    else:
        with pytest.raises(sb.ValueError_) as e:  # backend.rs:214-219: registered but unavailable
            sb.with_backend("cuda_fft", lambda: 0)
        assert "not available" in str(e.value)
    with pytest.raises(sb.ValueError_):
        sb.with_backend("nope", lambda: 0)
    with sb.fft_context().workers(2).cache_enabled(False).build():
        assert not c.is_enabled()
    assert c.is_enabled()
    with pytest.raises(sb.ValueError_):
        sb.PlanBuilder().build()
    assert sb.PlannerBackend.CUDA.value == "cuda"
