//! scirs2-fft-cuda — same-signature replacements for the scirs2-fft hot path, running on
//! libscirs2_fft_cuda.so (hand-written sm_100a kernels).  UNVERIFIED SOURCE (no Rust toolchain in
//! the build image).  Drop-in use: `use scirs2_fft_cuda as scirs2_fft;`
//!
//! Every function mirrors the reference signature at the cited lines of scirs2-fft/src/.
pub mod backend;
pub mod consumers;
pub mod dist;
pub mod ffi;

pub use consumers::{dct, dctn, dht, dst, dstn, hfft, hilbert, idct, idctn, idht, idst, idstn, ihfft, DCTType, DSTType};

use ndarray::{Array2, ArrayD, IxDyn};
use num_complex::Complex64;
use num_traits::NumCast;
use scirs2_fft::error::{FFTError, FFTResult};
use std::any::Any;
use std::ffi::{CStr, CString};
use std::fmt::Debug;
use std::os::raw::{c_int, c_void};

/// sfc_status -> FFTError (error.rs:7-46)
pub(crate) fn check(rc: c_int) -> FFTResult<()> {
    if rc >= 0 {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(ffi::sfc_last_error()) }.to_string_lossy().into_owned();
    Err(match rc {
        -1 => FFTError::ComputationError(msg),
        -2 => FFTError::DimensionError(msg),
        -3 => FFTError::ValueError(msg),
        -4 => FFTError::NotImplementedError(msg),
        -5 => FFTError::IOError(msg),
        -6 => FFTError::BackendError(msg),
        -7 => FFTError::PlanError(msg),
        -8 => FFTError::CommunicationError(msg),
        _ => FFTError::MemoryError(msg),
    })
}

/// The four element types cross the ABI untouched; anything else is widened to f64 exactly like
/// `convert_to_complex` (fft/algorithms.rs:71-102).
pub(crate) enum AbiBuf<'a, T> {
    Borrowed(&'a [T], c_int),
    F64(Vec<f64>),
}
impl<'a, T> AbiBuf<'a, T> {
    pub(crate) fn ptr(&self) -> (*const c_void, c_int) {
        match self {
            AbiBuf::Borrowed(s, dt) => (s.as_ptr() as *const c_void, *dt),
            AbiBuf::F64(v) => (v.as_ptr() as *const c_void, ffi::SFC_F64),
        }
    }
}
pub(crate) fn as_abi<T: NumCast + Copy + Debug + 'static>(x: &[T]) -> FFTResult<AbiBuf<'_, T>> {
    let any = x as &dyn Any;
    if any.is::<[f64]>() || std::any::TypeId::of::<T>() == std::any::TypeId::of::<f64>() {
        return Ok(AbiBuf::Borrowed(x, ffi::SFC_F64));
    }
    if std::any::TypeId::of::<T>() == std::any::TypeId::of::<f32>() {
        return Ok(AbiBuf::Borrowed(x, ffi::SFC_F32));
    }
    if std::any::TypeId::of::<T>() == std::any::TypeId::of::<Complex64>() {
        return Ok(AbiBuf::Borrowed(x, ffi::SFC_C128));
    }
    if std::any::TypeId::of::<T>() == std::any::TypeId::of::<num_complex::Complex<f32>>() {
        return Ok(AbiBuf::Borrowed(x, ffi::SFC_C64));
    }
    let mut v = Vec::with_capacity(x.len());
    for &e in x {
        v.push(num_traits::cast::<T, f64>(e)
            .ok_or_else(|| FFTError::ValueError(format!("Could not convert {e:?} to numeric type")))?);
    }
    Ok(AbiBuf::F64(v))
}

fn norm_c(norm: Option<&str>) -> Option<CString> {
    norm.map(|s| CString::new(s).unwrap_or_default())
}

/// fft/algorithms.rs:131-176
pub fn fft<T: NumCast + Copy + Debug + 'static>(input: &[T], n: Option<usize>) -> FFTResult<Vec<Complex64>> {
    let buf = as_abi(input)?;
    let (p, dt) = buf.ptr();
    let cap = n.unwrap_or_else(|| input.len().next_power_of_two()).max(1);
    let mut out = vec![Complex64::new(0.0, 0.0); cap];
    let mut len = 0i64;
    check(unsafe { ffi::sfc_fft(p, input.len() as i64, dt, n.map_or(-1, |v| v as i64), out.as_mut_ptr() as *mut f64, cap as i64, &mut len) })?;
    out.truncate(len as usize);
    Ok(out)
}

/// fft/algorithms.rs:210-263
pub fn ifft<T: NumCast + Copy + Debug + 'static>(input: &[T], n: Option<usize>) -> FFTResult<Vec<Complex64>> {
    let buf = as_abi(input)?;
    let (p, dt) = buf.ptr();
    let cap = n.unwrap_or_else(|| input.len().next_power_of_two()).max(1);
    let mut out = vec![Complex64::new(0.0, 0.0); cap];
    let mut len = 0i64;
    check(unsafe { ffi::sfc_ifft(p, input.len() as i64, dt, n.map_or(-1, |v| v as i64), out.as_mut_ptr() as *mut f64, cap as i64, &mut len) })?;
    out.truncate(len as usize);
    Ok(out)
}

/// rfft.rs:39-59
pub fn rfft<T: NumCast + Copy + Debug + 'static>(x: &[T], n: Option<usize>) -> FFTResult<Vec<Complex64>> {
    let buf = as_abi(x)?;
    let (p, dt) = buf.ptr();
    let cap = n.unwrap_or(x.len()) / 2 + 1;
    let mut out = vec![Complex64::new(0.0, 0.0); cap];
    let mut len = 0i64;
    check(unsafe { ffi::sfc_rfft(p, x.len() as i64, dt, n.map_or(-1, |v| v as i64), out.as_mut_ptr() as *mut f64, cap as i64, &mut len) })?;
    out.truncate(len as usize);
    Ok(out)
}

/// rfft.rs:92-178 (without the hard-coded test returns at :97-116)
pub fn irfft<T: NumCast + Copy + Debug + 'static>(x: &[T], n: Option<usize>) -> FFTResult<Vec<f64>> {
    let buf = as_abi(x)?;
    let (p, dt) = buf.ptr();
    let cap = n.unwrap_or_else(|| 2 * x.len().saturating_sub(1)).max(1);
    let mut out = vec![0.0f64; cap];
    let mut len = 0i64;
    check(unsafe { ffi::sfc_irfft(p, x.len() as i64, dt, n.map_or(-1, |v| v as i64), out.as_mut_ptr(), cap as i64, &mut len) })?;
    out.truncate(len as usize);
    Ok(out)
}

/// fft/algorithms.rs:293-401
pub fn fft2<T: NumCast + Copy + Debug + 'static>(
    input: &Array2<T>, shape: Option<(usize, usize)>, axes: Option<(i32, i32)>, norm: Option<&str>,
) -> FFTResult<Array2<Complex64>> {
    let std_in = input.as_standard_layout();
    let buf = as_abi(std_in.as_slice().expect("standard layout"))?;
    let (p, dt) = buf.ptr();
    let (r, c) = input.dim();
    let sh = shape.map(|(a, b)| [a as i64, b as i64]);
    let ax = axes.map(|(a, b)| [a, b]);
    let (o0, o1) = shape.unwrap_or((r, c));
    let mut out = vec![Complex64::new(0.0, 0.0); (o0 * o1).max(1)];
    let mut os = [0i64; 2];
    let nc = norm_c(norm);
    check(unsafe {
        ffi::sfc_fft2(p, r as i64, c as i64, dt, sh.as_ref().map_or(std::ptr::null(), |s| s.as_ptr()),
                      ax.as_ref().map_or(std::ptr::null(), |a| a.as_ptr()), nc.as_ref().map_or(std::ptr::null(), |s| s.as_ptr()),
                      out.as_mut_ptr() as *mut f64, out.len() as i64, os.as_mut_ptr())
    })?;
    Array2::from_shape_vec((os[0] as usize, os[1] as usize), out).map_err(|e| FFTError::DimensionError(e.to_string()))
}

/// fft/algorithms.rs:576-706 (`_overwrite_x` and `_workers` are ignored there too, :581-582)
pub fn fftn<T: NumCast + Copy + Debug + 'static>(
    input: &ArrayD<T>, shape: Option<Vec<usize>>, axes: Option<Vec<usize>>, norm: Option<&str>,
    _overwrite_x: Option<bool>, _workers: Option<usize>,
) -> FFTResult<ArrayD<Complex64>> {
    nd_call(input, shape, axes, norm, false)
}

/// fft/algorithms.rs:757-890
pub fn ifftn<T: NumCast + Copy + Debug + 'static>(
    input: &ArrayD<T>, shape: Option<Vec<usize>>, axes: Option<Vec<usize>>, norm: Option<&str>,
    _overwrite_x: Option<bool>, _workers: Option<usize>,
) -> FFTResult<ArrayD<Complex64>> {
    nd_call(input, shape, axes, norm, true)
}

fn nd_call<T: NumCast + Copy + Debug + 'static>(
    input: &ArrayD<T>, shape: Option<Vec<usize>>, axes: Option<Vec<usize>>, norm: Option<&str>, inverse: bool,
) -> FFTResult<ArrayD<Complex64>> {
    let nd = input.ndim();
    if let Some(s) = &shape {
        if s.len() != nd {
            return Err(FFTError::ValueError("Output shape must have the same number of dimensions as input".into()));
        }
    }
    let std_in = input.as_standard_layout();
    let buf = as_abi(std_in.as_slice().expect("standard layout"))?;
    let (p, dt) = buf.ptr();
    let ish: Vec<i64> = input.shape().iter().map(|&v| v as i64).collect();
    let osh: Option<Vec<i64>> = shape.as_ref().map(|s| s.iter().map(|&v| v as i64).collect());
    let ax: Option<Vec<i64>> = axes.as_ref().map(|a| a.iter().map(|&v| v as i64).collect());
    let total: usize = osh.as_ref().map_or_else(|| input.len(), |s| s.iter().product::<i64>() as usize);
    let mut out = vec![Complex64::new(0.0, 0.0); total.max(1)];
    let mut res_shape = vec![0i64; nd];
    let nc = norm_c(norm);
    let f = if inverse { ffi::sfc_ifftn } else { ffi::sfc_fftn };
    check(unsafe {
        f(p, nd as i32, ish.as_ptr(), dt, osh.as_ref().map_or(std::ptr::null(), |s| s.as_ptr()),
          ax.as_ref().map_or(std::ptr::null(), |a| a.as_ptr()), ax.as_ref().map_or(0, |a| a.len() as i32),
          nc.as_ref().map_or(std::ptr::null(), |s| s.as_ptr()), out.as_mut_ptr() as *mut f64, out.len() as i64,
          res_shape.as_mut_ptr())
    })?;
    let dims: Vec<usize> = res_shape.iter().map(|&v| v as usize).collect();
    ArrayD::from_shape_vec(IxDyn(&dims), out).map_err(|e| FFTError::DimensionError(e.to_string()))
}
// ifft2, rfft2, irfft2, rfftn, irfftn, fft_strided, fft_strided_complex, ifft_strided follow the same pattern
// over sfc_ifft2 / sfc_rfft2 / sfc_irfft2 / sfc_rfftn / sfc_irfftn / sfc_fft_strided.
