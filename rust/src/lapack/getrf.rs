//! `lapack::getrf` -- signature identical to the reference (src/lapack/getrf.rs:11-16); the
//! body is an FFI hop to the CUDA library instead of the scalar loops.
use std::any::TypeId;
use std::cmp;

use ndarray::{ArrayViewMut2, Axis};
use num_complex::Complex;

use crate::{ffi, Real, Scalar};

#[must_use]
pub fn getrf<A>(mut a: ArrayViewMut2<A>) -> (Vec<usize>, Option<usize>)
where
    A: Scalar,
    A::Real: Real,
{
    let (m, n) = (a.nrows() as i64, a.ncols() as i64);
    let dim_min = cmp::min(a.nrows(), a.ncols());
    let (rs, cs) = (a.stride_of(Axis(0)) as i64, a.stride_of(Axis(1)) as i64);
    let mut ipiv = vec![0_i64; dim_min];
    let mut info = -1_i64;
    let ptr = a.as_mut_ptr();
    let t = TypeId::of::<A>();
    // `A: 'static` follows from the `ScalarOperand` bound of `Scalar` (src/scalar.rs:344).
    let status = unsafe {
        if t == TypeId::of::<f64>() {
            ffi::lair_b200_dgetrf(m, n, ptr.cast(), rs, cs, ipiv.as_mut_ptr(), &mut info)
        } else if t == TypeId::of::<f32>() {
            ffi::lair_b200_sgetrf(m, n, ptr.cast(), rs, cs, ipiv.as_mut_ptr(), &mut info)
        } else if t == TypeId::of::<Complex<f64>>() {
            ffi::lair_b200_zgetrf(m, n, ptr.cast(), rs, cs, ipiv.as_mut_ptr(), &mut info)
        } else if t == TypeId::of::<Complex<f32>>() {
            ffi::lair_b200_cgetrf(m, n, ptr.cast(), rs, cs, ipiv.as_mut_ptr(), &mut info)
        } else {
            panic!("lair_b200: unsupported scalar type (f32, f64, Complex<f32>, Complex<f64> only; no CPU fallback)")
        }
    };
    ffi::check(status);
    let pivots = ipiv.into_iter().map(|p| p as usize).collect();
    let singular = if info < 0 { None } else { Some(info as usize) };
    (pivots, singular)
}
