//! `impl FftBackend` for the CUDA library (scirs2-fft/src/backend.rs:14-48) and its registration
//! (`BackendManager::register_backend`, :184-194).  UNVERIFIED SOURCE.
use crate::{check, ffi};
use num_complex::Complex64;
use scirs2_fft::backend::{get_backend_manager, FftBackend};
use scirs2_fft::error::FFTResult;
use std::ffi::CString;
use std::sync::Arc;

pub struct CudaFftBackend;

impl FftBackend for CudaFftBackend {
    fn name(&self) -> &str {
        "cuda_fft" // the id the reference's own example uses (examples/backend_example.rs:103-104)
    }
    fn description(&self) -> &str {
        "B200-native CUDA FFT (sm_100a Stockham tile kernels, four-step, Bluestein)"
    }
    fn is_available(&self) -> bool {
        unsafe { ffi::sfc_is_available() != 0 }
    }
    fn fft(&self, input: &[Complex64], output: &mut [Complex64]) -> FFTResult<()> {
        self.fft_sized(input, output, input.len())
    }
    fn ifft(&self, input: &[Complex64], output: &mut [Complex64]) -> FFTResult<()> {
        self.ifft_sized(input, output, input.len())
    }
    fn fft_sized(&self, input: &[Complex64], output: &mut [Complex64], size: usize) -> FFTResult<()> {
        // the size check and its message live in the library (backend.rs:96-100)
        check(unsafe {
            ffi::sfc_backend_fft_sized(input.as_ptr() as *const f64, input.len() as i64, output.as_mut_ptr() as *mut f64,
                                       output.len() as i64, size as i64)
        })
    }
    fn ifft_sized(&self, input: &[Complex64], output: &mut [Complex64], size: usize) -> FFTResult<()> {
        check(unsafe {
            ffi::sfc_backend_ifft_sized(input.as_ptr() as *const f64, input.len() as i64, output.as_mut_ptr() as *mut f64,
                                        output.len() as i64, size as i64)
        })
    }
    fn supports_feature(&self, feature: &str) -> bool {
        CString::new(feature).map_or(false, |c| unsafe { ffi::sfc_backend_supports_feature(c.as_ptr()) != 0 })
    }
}

/// Register the backend under "cuda_fft" and make it current.
pub fn install() -> FFTResult<()> {
    let m = get_backend_manager();
    m.register_backend("cuda_fft".to_string(), Arc::new(CudaFftBackend))?;
    m.set_backend("cuda_fft")
}
