"""The header-only C++ mirror (include/scirs2_fft_cuda.hpp) compiles against the C ABI and behaves
like the reference's own tests; without a GPU it must fail loudly with FFTError::Backend."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        from scirs_b200 import _lib

        return _lib.load().sfc_device_count() > 0
    except Exception:
        return False


def _build(build_artifacts):
    exe = os.path.join(ROOT, "build", "cpp_mirror_test")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    libdir = os.path.join(ROOT, "scirs_b200", "lib")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp_mirror_test.cpp"), "-o", exe, "-L", libdir,
                    "-lscirs2_fft_cuda", f"-Wl,-rpath,{libdir}"], check=True)
    return exe


@pytest.mark.skipif(_has_gpu(), reason="CPU-only behaviour")
def test_cpp_mirror_fails_loudly_without_gpu(build_artifacts):
    out = subprocess.run([_build(build_artifacts)], capture_output=True, text=True)
    assert out.returncode == 0 and "BackendError ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_cpp_mirror_on_gpu(build_artifacts):
    out = subprocess.run([_build(build_artifacts)], capture_output=True, text=True)
    assert out.returncode == 0 and "cpp mirror ok" in out.stdout, out.stdout + out.stderr
