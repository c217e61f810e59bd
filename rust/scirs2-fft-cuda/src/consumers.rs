//! Same-signature replacements for the in-crate consumers of the FFT path: dct.rs:56-420, dst.rs:48-405,
//! hartley.rs:37-130, hfft/complex_to_real.rs:58-135, hfft/real_to_complex.rs:49-149, lib.rs:437-516.
//! UNVERIFIED SOURCE (no Rust toolchain in the build image); the same entry points are exercised from Python
//! (scirs_b200/consumers.py, tests/test_gpu_consumers.py) and C++ (include/scirs2_fft_cuda.hpp).
use crate::{as_abi, check, ffi};
use ndarray::{ArrayD, ArrayView, IxDyn};
use num_complex::Complex64;
use num_traits::NumCast;
use scirs2_fft::error::{FFTError, FFTResult};
use std::ffi::CString;
use std::fmt::Debug;

/// dct.rs:13-23
#[derive(Debug, Copy, Clone, PartialEq, Eq)]
pub enum DCTType {
    Type1 = 1,
    Type2 = 2,
    Type3 = 3,
    Type4 = 4,
}
/// dst.rs:13-22
#[derive(Debug, Copy, Clone, PartialEq, Eq)]
pub enum DSTType {
    Type1 = 1,
    Type2 = 2,
    Type3 = 3,
    Type4 = 4,
}

fn widen<T: NumCast + Copy + Debug>(x: &[T]) -> FFTResult<Vec<f64>> {
    x.iter()
        .map(|&v| num_traits::cast::<T, f64>(v).ok_or_else(|| FFTError::ValueError(format!("Could not convert {v:?} to f64"))))
        .collect()
}

type TrigFn = unsafe extern "C" fn(*const f64, i32, *const i64, *const i32, i32, i32, i32, *const std::os::raw::c_char, *mut f64) -> i32;

fn trig(f: TrigFn, data: &[f64], shape: &[i64], axes: Option<&[i32]>, ttype: i32, inverse: bool, norm: Option<&str>) -> FFTResult<Vec<f64>> {
    let mut out = vec![0.0f64; data.len().max(1)];
    let nc = norm.map(|s| CString::new(s).unwrap_or_default());
    check(unsafe {
        f(data.as_ptr(), shape.len() as i32, shape.as_ptr(), axes.map_or(std::ptr::null(), |a| a.as_ptr()),
          axes.map_or(0, |a| a.len() as i32), ttype, inverse as i32, nc.as_ref().map_or(std::ptr::null(), |c| c.as_ptr()),
          out.as_mut_ptr())
    })?;
    out.truncate(data.len());
    Ok(out)
}

/// dct.rs:56-78
pub fn dct<T: NumCast + Copy + Debug>(x: &[T], dct_type: Option<DCTType>, norm: Option<&str>) -> FFTResult<Vec<f64>> {
    let v = widen(x)?;
    trig(ffi::sfc_dct, &v, &[v.len() as i64], Some(&[0]), dct_type.unwrap_or(DCTType::Type2) as i32, false, norm)
}
/// dct.rs:115-138
pub fn idct<T: NumCast + Copy + Debug>(x: &[T], dct_type: Option<DCTType>, norm: Option<&str>) -> FFTResult<Vec<f64>> {
    let v = widen(x)?;
    trig(ffi::sfc_dct, &v, &[v.len() as i64], Some(&[0]), dct_type.unwrap_or(DCTType::Type2) as i32, true, norm)
}
/// dst.rs:48-70
pub fn dst<T: NumCast + Copy + Debug>(x: &[T], dst_type: Option<DSTType>, norm: Option<&str>) -> FFTResult<Vec<f64>> {
    let v = widen(x)?;
    trig(ffi::sfc_dst, &v, &[v.len() as i64], Some(&[0]), dst_type.unwrap_or(DSTType::Type2) as i32, false, norm)
}
/// dst.rs:103-126
pub fn idst<T: NumCast + Copy + Debug>(x: &[T], dst_type: Option<DSTType>, norm: Option<&str>) -> FFTResult<Vec<f64>> {
    let v = widen(x)?;
    trig(ffi::sfc_dst, &v, &[v.len() as i64], Some(&[0]), dst_type.unwrap_or(DSTType::Type2) as i32, true, norm)
}

fn trig_nd<T: NumCast + Copy + Debug>(f: TrigFn, x: &ArrayView<T, IxDyn>, ttype: i32, inverse: bool, norm: Option<&str>,
                                      axes: Option<Vec<usize>>) -> FFTResult<ArrayD<f64>> {
    let std_in = x.as_standard_layout();
    let v = widen(std_in.as_slice().expect("standard layout"))?;
    let shape: Vec<i64> = x.shape().iter().map(|&s| s as i64).collect();
    let ax: Option<Vec<i32>> = axes.map(|a| a.iter().map(|&i| i as i32).collect());
    let out = trig(f, &v, &shape, ax.as_deref(), ttype, inverse, norm)?;
    ArrayD::from_shape_vec(IxDyn(x.shape()), out).map_err(|e| FFTError::DimensionError(e.to_string()))
}
/// dct.rs:302-360 (dct2 = the same call with axes [1, 0], dct.rs:168-206)
pub fn dctn<T: NumCast + Copy + Debug>(x: &ArrayView<T, IxDyn>, dct_type: Option<DCTType>, norm: Option<&str>,
                                       axes: Option<Vec<usize>>) -> FFTResult<ArrayD<f64>> {
    trig_nd(ffi::sfc_dct, x, dct_type.unwrap_or(DCTType::Type2) as i32, false, norm, axes)
}
/// dct.rs:373-420
pub fn idctn<T: NumCast + Copy + Debug>(x: &ArrayView<T, IxDyn>, dct_type: Option<DCTType>, norm: Option<&str>,
                                        axes: Option<Vec<usize>>) -> FFTResult<ArrayD<f64>> {
    trig_nd(ffi::sfc_dct, x, dct_type.unwrap_or(DCTType::Type2) as i32, true, norm, axes)
}
/// dst.rs:284-340
pub fn dstn<T: NumCast + Copy + Debug>(x: &ArrayView<T, IxDyn>, dst_type: Option<DSTType>, norm: Option<&str>,
                                       axes: Option<Vec<usize>>) -> FFTResult<ArrayD<f64>> {
    trig_nd(ffi::sfc_dst, x, dst_type.unwrap_or(DSTType::Type2) as i32, false, norm, axes)
}
/// dst.rs:354-405
pub fn idstn<T: NumCast + Copy + Debug>(x: &ArrayView<T, IxDyn>, dst_type: Option<DSTType>, norm: Option<&str>,
                                        axes: Option<Vec<usize>>) -> FFTResult<ArrayD<f64>> {
    trig_nd(ffi::sfc_dst, x, dst_type.unwrap_or(DSTType::Type2) as i32, true, norm, axes)
}

/// hartley.rs:37-66 (any shape is flattened there too)
pub fn dht(x: &[f64]) -> FFTResult<Vec<f64>> {
    let mut out = vec![0.0; x.len().max(1)];
    check(unsafe { ffi::sfc_dht(x.as_ptr(), x.len() as i64, out.as_mut_ptr()) })?;
    out.truncate(x.len());
    Ok(out)
}
/// hartley.rs:92-112
pub fn idht(h: &[f64]) -> FFTResult<Vec<f64>> {
    let mut out = vec![0.0; h.len().max(1)];
    check(unsafe { ffi::sfc_idht(h.as_ptr(), h.len() as i64, out.as_mut_ptr()) })?;
    out.truncate(h.len());
    Ok(out)
}

/// hfft/complex_to_real.rs:58-135 (`norm` is ignored there too)
pub fn hfft<T: NumCast + Copy + Debug + 'static>(x: &[T], n: Option<usize>, _norm: Option<&str>) -> FFTResult<Vec<f64>> {
    let buf = as_abi(x)?;
    let (p, dt) = buf.ptr();
    let cap = n.unwrap_or(x.len()).max(1);
    let mut out = vec![0.0; cap];
    let mut len = 0i64;
    check(unsafe { ffi::sfc_hfft(p, x.len() as i64, dt, n.map_or(-1, |v| v as i64), out.as_mut_ptr(), cap as i64, &mut len) })?;
    out.truncate(len as usize);
    Ok(out)
}
/// hfft/real_to_complex.rs:49-149
pub fn ihfft<T: NumCast + Copy + Debug>(x: &[T], n: Option<usize>, _norm: Option<&str>) -> FFTResult<Vec<Complex64>> {
    let v = widen(x)?;
    let cap = n.unwrap_or(v.len()).max(1);
    let mut out = vec![Complex64::new(0.0, 0.0); cap];
    let mut len = 0i64;
    check(unsafe { ffi::sfc_ihfft(v.as_ptr(), v.len() as i64, n.map_or(-1, |k| k as i64), out.as_mut_ptr() as *mut f64, cap as i64, &mut len) })?;
    out.truncate(len as usize);
    Ok(out)
}
/// lib.rs:437-516
pub fn hilbert<T: NumCast + Copy + Debug>(x: &[T]) -> FFTResult<Vec<Complex64>> {
    let v = widen(x)?;
    let mut out = vec![Complex64::new(0.0, 0.0); v.len().max(1)];
    check(unsafe { ffi::sfc_hilbert(v.as_ptr(), v.len() as i64, out.as_mut_ptr() as *mut f64) })?;
    out.truncate(v.len());
    Ok(out)
}
