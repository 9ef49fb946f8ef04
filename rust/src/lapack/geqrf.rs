//! `lapack::geqrf` -- signature identical to the reference (src/lapack/geqrf.rs:9-13); the body is an FFI hop to the
//! CUDA library (cluster panel + compact-WY update for f32 / f64, one reflector at a time for the complex types).
use std::any::TypeId;
use std::cmp;
use std::ops::{Div, MulAssign};

use ndarray::{Array1, ArrayBase, Axis, DataMut, Ix2};
use num_complex::Complex;

use crate::{ffi, Scalar};

/// Computes the QR factorization of a matrix.
pub fn geqrf<A, S>(a: &mut ArrayBase<S, Ix2>) -> Array1<A>
where
    A: Scalar + Div<<A as Scalar>::Real, Output = A> + MulAssign<<A as Scalar>::Real>,
    S: DataMut<Elem = A>,
{
    let (m, n) = (a.nrows() as i64, a.ncols() as i64);
    let min_dim = cmp::min(a.nrows(), a.ncols());
    let mut tau = Array1::<A>::zeros(min_dim);
    if min_dim == 0 {
        return tau;
    }
    let (rs, cs) = (a.stride_of(Axis(0)) as i64, a.stride_of(Axis(1)) as i64);
    let (ptr, tptr) = (a.as_mut_ptr(), tau.as_mut_ptr());
    let t = TypeId::of::<A>();
    let status = unsafe {
        if t == TypeId::of::<f64>() {
            ffi::lair_b200_dgeqrf(m, n, ptr.cast(), rs, cs, tptr.cast())
        } else if t == TypeId::of::<f32>() {
            ffi::lair_b200_sgeqrf(m, n, ptr.cast(), rs, cs, tptr.cast())
        } else if t == TypeId::of::<Complex<f64>>() {
            ffi::lair_b200_zgeqrf(m, n, ptr.cast(), rs, cs, tptr.cast())
        } else if t == TypeId::of::<Complex<f32>>() {
            ffi::lair_b200_cgeqrf(m, n, ptr.cast(), rs, cs, tptr.cast())
        } else {
            panic!("lair_b200: unsupported scalar type (f32, f64, Complex<f32>, Complex<f64> only; no CPU fallback)")
        }
    };
    ffi::check(status);
    tau
}

/// `qr::Factorized::q` (src/decomposition/qr.rs:27-59) as one call: the nrows x nrows unitary factor from the factored
/// matrix and tau.  `qr::Factorized::r` stays as in the reference (it only zeroes the strict lower triangle).
pub fn qr_q<A, S>(qr: &ArrayBase<S, Ix2>, tau: &Array1<A>) -> ndarray::Array2<A>
where
    A: Scalar,
    S: ndarray::Data<Elem = A>,
{
    let (m, n) = (qr.nrows() as i64, qr.ncols() as i64);
    let mut q = ndarray::Array2::<A>::zeros((qr.nrows(), qr.nrows()));
    if m == 0 {
        return q;
    }
    let (rs, cs) = (qr.stride_of(Axis(0)) as i64, qr.stride_of(Axis(1)) as i64);
    let t = TypeId::of::<A>();
    let status = unsafe {
        if t == TypeId::of::<f64>() {
            ffi::lair_b200_dqr_q(m, n, qr.as_ptr().cast(), rs, cs, tau.as_ptr().cast(), q.as_mut_ptr().cast(), m, 1)
        } else if t == TypeId::of::<f32>() {
            ffi::lair_b200_sqr_q(m, n, qr.as_ptr().cast(), rs, cs, tau.as_ptr().cast(), q.as_mut_ptr().cast(), m, 1)
        } else if t == TypeId::of::<Complex<f64>>() {
            ffi::lair_b200_zqr_q(m, n, qr.as_ptr().cast(), rs, cs, tau.as_ptr().cast(), q.as_mut_ptr().cast(), m, 1)
        } else if t == TypeId::of::<Complex<f32>>() {
            ffi::lair_b200_cqr_q(m, n, qr.as_ptr().cast(), rs, cs, tau.as_ptr().cast(), q.as_mut_ptr().cast(), m, 1)
        } else {
            panic!("lair_b200: unsupported scalar type (f32, f64, Complex<f32>, Complex<f64> only; no CPU fallback)")
        }
    };
    ffi::check(status);
    q
}
