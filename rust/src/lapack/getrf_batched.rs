//! Batched LU of independent small matrices (BASELINE configs[2]; SURVEY 8e "batched small LU").
//! The reference has no batched entry point: a caller loops `getrf` (src/lapack/getrf.rs:12-27)
//! over the matrices.  This is that loop as ONE call on a standard-layout `[batch][n][n]` array
//! (n <= 32); every matrix comes back bit-identical to the reference's row-major body
//! (getrf.rs:46-120), with the same (pivots, last zero-pivot step) per matrix.
use std::any::TypeId;

use ndarray::ArrayViewMut3;

use crate::{ffi, Real, Scalar};

/// Returns, per matrix, what `getrf` returns for it.
#[must_use]
pub fn getrf_batched<A>(mut a: ArrayViewMut3<A>) -> Vec<(Vec<usize>, Option<usize>)>
where
    A: Scalar,
    A::Real: Real,
{
    let (batch, n, n2) = a.dim();
    assert_eq!(n, n2, "batched LU takes square matrices");
    assert!(n <= 32, "batched LU takes matrices of order <= 32; use getrf");
    assert!(a.is_standard_layout(), "batched LU takes a standard-layout [batch][n][n] array");
    let mut ipiv = vec![0_i32; batch * n];
    let mut info = vec![-1_i32; batch];
    let t = TypeId::of::<A>();
    let status = unsafe {
        if t == TypeId::of::<f64>() {
            ffi::lair_b200_dgetrf_batched(batch as i64, n as i64, a.as_mut_ptr().cast(), ipiv.as_mut_ptr(), info.as_mut_ptr())
        } else if t == TypeId::of::<f32>() {
            ffi::lair_b200_sgetrf_batched(batch as i64, n as i64, a.as_mut_ptr().cast(), ipiv.as_mut_ptr(), info.as_mut_ptr())
        } else {
            panic!("lair_b200: batched LU supports f32 and f64 only (no CPU fallback)")
        }
    };
    ffi::check(status);
    (0..batch)
        .map(|i| {
            let pivots = ipiv[i * n..(i + 1) * n].iter().map(|&p| p as usize).collect();
            let singular = if info[i] < 0 { None } else { Some(info[i] as usize) };
            (pivots, singular)
        })
        .collect()
}
