//! `equation::solve` -- signature, validation order and error messages of the reference
//! (src/equation.rs:32-60).  For f32 / f64 the copy of `a`, the factorization and the
//! substitution are ONE FFI call (`lair_b200_{s,d}gesv`): `a` is uploaded once (the upload is
//! the reference's `a.to_owned()`), L\U and the pivots never leave HBM, only `x` comes back.
//! Complex scalars take the reference's own route through `lu::Factorized`, whose `From` and
//! `solve` already run on the device (decomposition/lu_device.rs).
//! Source-only like the rest of this directory; the same entry point is exercised from Python by
//! tests/test_gpu_parity.py (`lair_b200.equation.solve`) and timed by bench.py's `e2e` leg.
use std::any::TypeId;

use ndarray::{Array1, ArrayBase, Axis, Data, Ix1, Ix2};

use crate::decomposition::lu;
use crate::{ffi, InvalidInput, Real, Scalar};

pub fn solve<A, SA, SB>(a: &ArrayBase<SA, Ix2>, b: &ArrayBase<SB, Ix1>) -> Result<Array1<A>, InvalidInput>
where
    A: Scalar,
    A::Real: Real,
    SA: Data<Elem = A>,
    SB: Data<Elem = A>,
{
    if a.nrows() != a.ncols() {
        return Err(InvalidInput::Shape("input matrix is not square".to_string()));
    }
    if a.nrows() != b.len() {
        return Err(InvalidInput::Shape(format!(
            "The number of elements in `b`, {}, must be the same as the number of rows in `a`, {}",
            b.len(),
            a.nrows()
        )));
    }
    let t = TypeId::of::<A>();
    let real = t == TypeId::of::<f64>() || t == TypeId::of::<f32>();
    if !real {
        let factorized = lu::Factorized::from(a.to_owned());
        return if factorized.is_singular() {
            Err(InvalidInput::Value("`a` is a singular matrix".to_string()))
        } else {
            factorized.solve(b)
        };
    }
    let n = a.nrows() as i64;
    let (ars, acs) = (a.stride_of(Axis(0)) as i64, a.stride_of(Axis(1)) as i64);
    let brs = b.stride_of(Axis(0)) as i64;
    let mut x = Array1::<A>::zeros(b.len());
    let mut info = -1_i64;
    let status = unsafe {
        if t == TypeId::of::<f64>() {
            ffi::lair_b200_dgesv(n, 1, a.as_ptr().cast(), ars, acs, b.as_ptr().cast(), brs, 1,
                                 x.as_mut_ptr().cast(), 1, 1, &mut info)
        } else {
            ffi::lair_b200_sgesv(n, 1, a.as_ptr().cast(), ars, acs, b.as_ptr().cast(), brs, 1,
                                 x.as_mut_ptr().cast(), 1, 1, &mut info)
        }
    };
    ffi::check(status);
    if info >= 0 {
        // a zero pivot was met: the reference reports it before it ever substitutes (equation.rs:55-56)
        Err(InvalidInput::Value("`a` is a singular matrix".to_string()))
    } else {
        Ok(x)
    }
}
