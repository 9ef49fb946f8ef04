//! `lapack::getrs` -- signature and panics identical to the reference (src/lapack/getrs.rs:12-20).
use std::any::TypeId;

use ndarray::{Array1, ArrayBase, Axis, Data, Ix1, Ix2};
use num_complex::Complex;

use crate::{ffi, Scalar};

pub fn getrs<A, SA, SB>(a: &ArrayBase<SA, Ix2>, p: &[usize], b: &ArrayBase<SB, Ix1>) -> Array1<A>
where
    A: Scalar,
    SA: Data<Elem = A>,
    SB: Data<Elem = A>,
{
    assert_eq!(a.nrows(), p.len());
    assert_eq!(p.len(), b.len());
    assert!(a.ncols() >= p.len());
    let n = p.len() as i64;
    let ipiv: Vec<i64> = p.iter().map(|&v| v as i64).collect();
    let mut x = Array1::<A>::zeros(p.len());
    let (lrs, lcs) = (a.stride_of(Axis(0)) as i64, a.stride_of(Axis(1)) as i64);
    let brs = b.stride_of(Axis(0)) as i64;
    let t = TypeId::of::<A>();
    let status = unsafe {
        if t == TypeId::of::<f64>() {
            ffi::lair_b200_dgetrs(n, 1, a.as_ptr().cast(), lrs, lcs, ipiv.as_ptr(), b.as_ptr().cast(), brs, 1,
                                  x.as_mut_ptr().cast(), 1, 1)
        } else if t == TypeId::of::<f32>() {
            ffi::lair_b200_sgetrs(n, 1, a.as_ptr().cast(), lrs, lcs, ipiv.as_ptr(), b.as_ptr().cast(), brs, 1,
                                  x.as_mut_ptr().cast(), 1, 1)
        } else if t == TypeId::of::<Complex<f64>>() {
            ffi::lair_b200_zgetrs(n, 1, a.as_ptr().cast(), lrs, lcs, ipiv.as_ptr(), b.as_ptr().cast(), brs, 1,
                                  x.as_mut_ptr().cast(), 1, 1)
        } else if t == TypeId::of::<Complex<f32>>() {
            ffi::lair_b200_cgetrs(n, 1, a.as_ptr().cast(), lrs, lcs, ipiv.as_ptr(), b.as_ptr().cast(), brs, 1,
                                  x.as_mut_ptr().cast(), 1, 1)
        } else {
            panic!("lair_b200: unsupported scalar type (no CPU fallback)")
        }
    };
    ffi::check(status);
    x
}
