// Links the prebuilt C-ABI library; SCIRS2_FFT_CUDA_LIB_DIR points at the directory holding
// libscirs2_fft_cuda.so (scirs_b200/lib/ in this repository).
fn main() {
    let dir = std::env::var("SCIRS2_FFT_CUDA_LIB_DIR").unwrap_or_else(|_| "../../scirs_b200/lib".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=scirs2_fft_cuda");
    println!("cargo:rerun-if-env-changed=SCIRS2_FFT_CUDA_LIB_DIR");
}
