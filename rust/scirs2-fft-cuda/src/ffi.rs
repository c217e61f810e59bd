//! Raw bindings of include/scirs2_fft_cuda.h (ABI version 1).  UNVERIFIED: never compiled here.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_double, c_int, c_void};

pub const SFC_MAX_DIMS: usize = 8;
pub const SFC_F32: c_int = 0;
pub const SFC_F64: c_int = 1;
pub const SFC_C64: c_int = 2;
pub const SFC_C128: c_int = 3;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct sfc_desc {
    pub ndim: i32,
    pub shape: [i64; SFC_MAX_DIMS],
    pub naxes: i32,
    pub axes: [i32; SFC_MAX_DIMS],
    pub kind: i32,
    pub prec: i32,
    pub direction: i32,
    pub flags: i32,
    pub scale: c_double,
    pub in_shape: [i64; SFC_MAX_DIMS],
    pub scatter_parts: i32,
    pub reserved: i32,
    pub axis_in_len: i64,
    pub axis_out_len: i64,
    pub aux_in: *const core::ffi::c_void,
    pub aux_out: *const core::ffi::c_void,
    pub scale_dc: f64,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct sfc_cache_stats {
    pub hit_count: u64,
    pub miss_count: u64,
    pub hit_rate: f64,
    pub size: u64,
    pub max_size: u64,
}

pub enum sfc_plan {}

extern "C" {
    pub fn sfc_init(device: c_int) -> c_int;
    pub fn sfc_device_count() -> c_int;
    pub fn sfc_is_available() -> c_int;
    pub fn sfc_last_error() -> *const c_char;
    pub fn sfc_abi_version() -> c_int;

    pub fn sfc_plan_create(out: *mut *mut sfc_plan, desc: *const sfc_desc) -> c_int;
    pub fn sfc_plan_destroy(plan: *mut sfc_plan) -> c_int;
    pub fn sfc_exec_device(plan: *mut sfc_plan, d_in: *const c_void, d_out: *mut c_void, stream: *mut c_void) -> c_int;
    pub fn sfc_exec_host(plan: *mut sfc_plan, h_in: *const c_void, h_out: *mut c_void) -> c_int;

    pub fn sfc_cache_get_stats(out: *mut sfc_cache_stats) -> c_int;
    pub fn sfc_cache_set_enabled(enabled: c_int) -> c_int;
    pub fn sfc_cache_is_enabled() -> c_int;
    pub fn sfc_cache_clear() -> c_int;
    pub fn sfc_cache_configure(max_entries: u64, max_age_seconds: f64) -> c_int;

    pub fn sfc_fft(x: *const c_void, len: i64, dtype: c_int, n: i64, out: *mut f64, cap: i64, out_len: *mut i64) -> c_int;
    pub fn sfc_ifft(x: *const c_void, len: i64, dtype: c_int, n: i64, out: *mut f64, cap: i64, out_len: *mut i64) -> c_int;
    pub fn sfc_rfft(x: *const c_void, len: i64, dtype: c_int, n: i64, out: *mut f64, cap: i64, out_len: *mut i64) -> c_int;
    pub fn sfc_irfft(x: *const c_void, len: i64, dtype: c_int, n: i64, out: *mut f64, cap: i64, out_len: *mut i64) -> c_int;
    pub fn sfc_fft2(x: *const c_void, rows: i64, cols: i64, dtype: c_int, shape2: *const i64, axes2: *const i32,
                    norm: *const c_char, out: *mut f64, cap: i64, out_shape2: *mut i64) -> c_int;
    pub fn sfc_ifft2(x: *const c_void, rows: i64, cols: i64, dtype: c_int, shape2: *const i64, axes2: *const i32,
                     norm: *const c_char, out: *mut f64, cap: i64, out_shape2: *mut i64) -> c_int;
    pub fn sfc_rfft2(x: *const c_void, rows: i64, cols: i64, dtype: c_int, shape2: *const i64, out: *mut f64,
                     cap: i64, out_shape2: *mut i64) -> c_int;
    pub fn sfc_irfft2(x: *const c_void, rows: i64, cols: i64, dtype: c_int, shape2: *const i64, out: *mut f64,
                      cap: i64, out_shape2: *mut i64) -> c_int;
    pub fn sfc_fftn(x: *const c_void, ndim: i32, in_shape: *const i64, dtype: c_int, shape: *const i64,
                    axes: *const i64, naxes: i32, norm: *const c_char, out: *mut f64, cap: i64, out_shape: *mut i64) -> c_int;
    pub fn sfc_ifftn(x: *const c_void, ndim: i32, in_shape: *const i64, dtype: c_int, shape: *const i64,
                     axes: *const i64, naxes: i32, norm: *const c_char, out: *mut f64, cap: i64, out_shape: *mut i64) -> c_int;
    pub fn sfc_rfftn(x: *const c_void, ndim: i32, in_shape: *const i64, dtype: c_int, shape: *const i64,
                     axes: *const i64, naxes: i32, norm: *const c_char, out: *mut f64, cap: i64, out_shape: *mut i64) -> c_int;
    pub fn sfc_irfftn(x: *const c_void, ndim: i32, in_shape: *const i64, dtype: c_int, shape: *const i64, nshape: i32,
                      axes: *const i64, naxes: i32, norm: *const c_char, out: *mut f64, cap: i64, out_shape: *mut i64) -> c_int;
    pub fn sfc_fft_strided(x: *const c_void, ndim: i32, in_shape: *const i64, dtype: c_int, axis: i64, inverse: c_int,
                           out: *mut f64, cap: i64) -> c_int;

    pub fn sfc_backend_fft_sized(input: *const f64, in_len: i64, output: *mut f64, out_len: i64, size: i64) -> c_int;
    pub fn sfc_backend_ifft_sized(input: *const f64, in_len: i64, output: *mut f64, out_len: i64, size: i64) -> c_int;
    pub fn sfc_backend_supports_feature(feature: *const c_char) -> c_int;
    pub fn sfc_execute_batch(inputs: *const f64, outputs: *mut f64, count: i64, size: i64, inverse: c_int) -> c_int;

    // consumers of the hot path (dct.rs, dst.rs, hartley.rs, hfft/*.rs, lib.rs::hilbert, spectrogram.rs, memory_efficient.rs)
    pub fn sfc_dct(x: *const f64, ndim: i32, shape: *const i64, axes: *const i32, naxes: i32, ttype: i32, inverse: i32,
                   norm: *const c_char, out: *mut f64) -> c_int;
    pub fn sfc_dst(x: *const f64, ndim: i32, shape: *const i64, axes: *const i32, naxes: i32, ttype: i32, inverse: i32,
                   norm: *const c_char, out: *mut f64) -> c_int;
    pub fn sfc_dht(x: *const f64, n: i64, out: *mut f64) -> c_int;
    pub fn sfc_idht(h: *const f64, n: i64, out: *mut f64) -> c_int;
    pub fn sfc_dht2(x: *const f64, rows: i64, cols: i64, axis0: i32, axis1: i32, out: *mut f64) -> c_int;
    pub fn sfc_hfft(x: *const c_void, len: i64, dtype: c_int, n: i64, out: *mut f64, cap: i64, out_len: *mut i64) -> c_int;
    pub fn sfc_ihfft(x: *const f64, len: i64, n: i64, out: *mut f64, cap: i64, out_len: *mut i64) -> c_int;
    pub fn sfc_hilbert(x: *const f64, n: i64, out: *mut f64) -> c_int;
    pub fn sfc_stft(x: *const f64, len: i64, window: *const f64, nperseg: i64, noverlap: i64, nfft: i64, detrend: i32,
                    onesided: i32, boundary: i32, out_mode: i32, scale: f64, out: *mut c_void, cap: i64,
                    freq_len: *mut i64, frames: *mut i64) -> c_int;
    pub fn sfc_signal_spectra(x: *const f64, len: i64, window: *const f64, nperseg: i64, step: i64, frames: i64, p: i64,
                              detrend: i32, reduce: i32, bins: i64, scale: f64, out: *mut c_void) -> c_int;
    pub fn sfc_fft_inplace(input: *mut f64, n: i64, output: *mut f64, out_len: i64, inverse: i32, normalize: i32) -> c_int;
    pub fn sfc_fft2_efficient(x: *const c_void, rows: i64, cols: i64, dtype: c_int, out_rows: i64, out_cols: i64,
                              inverse: i32, normalize: i32, out: *mut f64) -> c_int;
    pub fn sfc_fft_streaming(x: *const c_void, len: i64, dtype: c_int, n: i64, inverse: i32, chunk: i64, out: *mut f64) -> c_int;
    pub fn sfc_fftn_optimized(x: *const f64, ndim: i32, shape: *const i64, axes: *const i32, naxes: i32, out: *mut f64) -> c_int;
}
