//! Optional replacement for the body of `decomposition::lu::Factorized` (src/decomposition/lu.rs:12-171)
//! that keeps the factors in HBM between `from` and `solve`.  The public
//! surface -- `From<ArrayBase<S, Ix2>>`, `p`, `l`, `u`, `is_singular`, `solve`, `into_pl` -- is the
//! reference's, type parameters included: `Factorized<A, S>` with `#[derive(Debug)]`-equivalent formatting (lu.rs:11-20);
//! only the private fields change from (lu, pivots, singular) to a device handle (`S` is kept as a marker so that
//! code naming `lu::Factorized<f64, OwnedRepr<f64>>` keeps compiling).
//! Source-only like the rest of this directory (no Rust toolchain in the build image); the same
//! entry points are exercised from Python by tests/test_gpu_parity.py::test_lu_handle_*.
use std::any::TypeId;
use std::fmt;
use std::marker::PhantomData;
use std::os::raw::c_void;

use ndarray::{Array1, Array2, ArrayBase, Axis, Data, DataMut, Ix1, Ix2};

use crate::{ffi, InvalidInput, Real, Scalar};

const VIEW_L: i32 = 0;
const VIEW_U: i32 = 1;
const VIEW_P: i32 = 2;
const VIEW_PL: i32 = 3;

/// LU decomposition factors, device-resident.
pub struct Factorized<A, S>
where
    A: fmt::Debug,
    S: Data<Elem = A>,
{
    handle: *mut c_void, // lair_b200_lu_t: owns L\U and the pivots in HBM
    rows: usize,
    cols: usize,
    singular: Option<usize>,
    _marker: PhantomData<(A, S)>,
}

/// The reference derives `Debug` (lu.rs:11); the factors live in HBM, so the shape, the handle and `singular` are shown.
impl<A, S> fmt::Debug for Factorized<A, S>
where
    A: fmt::Debug,
    S: Data<Elem = A>,
{
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        f.debug_struct("Factorized")
            .field("lu", &format_args!("<{} x {} on device, handle {:p}>", self.rows, self.cols, self.handle))
            .field("singular", &self.singular)
            .finish()
    }
}

impl<A, S> Drop for Factorized<A, S>
where
    A: fmt::Debug,
    S: Data<Elem = A>,
{
    fn drop(&mut self) {
        unsafe { ffi::lair_b200_lu_destroy(self.handle) };
    }
}

impl<A, S> From<ArrayBase<S, Ix2>> for Factorized<A, S>
where
    A: Scalar,
    A::Real: Real,
    S: DataMut<Elem = A>,
{
    /// lu.rs:156-171.  The array is consumed; its storage is only read (the upload is the copy).
    fn from(a: ArrayBase<S, Ix2>) -> Self {
        let (m, n) = (a.nrows() as i64, a.ncols() as i64);
        let (rs, cs) = (a.stride_of(Axis(0)) as i64, a.stride_of(Axis(1)) as i64);
        let mut handle: *mut c_void = std::ptr::null_mut();
        let mut info = -1_i64;
        let p = a.as_ptr();
        let t = TypeId::of::<A>();
        let status = unsafe {
            if t == TypeId::of::<f32>() {
                ffi::lair_b200_slu_factor(m, n, p.cast(), rs, cs, &mut handle, &mut info)
            } else if t == TypeId::of::<f64>() {
                ffi::lair_b200_dlu_factor(m, n, p.cast(), rs, cs, &mut handle, &mut info)
            } else if t == TypeId::of::<num_complex::Complex<f32>>() {
                ffi::lair_b200_clu_factor(m, n, p.cast(), rs, cs, &mut handle, &mut info)
            } else if t == TypeId::of::<num_complex::Complex<f64>>() {
                ffi::lair_b200_zlu_factor(m, n, p.cast(), rs, cs, &mut handle, &mut info)
            } else {
                panic!("lair_b200 implements f32, f64, Complex<f32>, Complex<f64> (no CPU fallback)")
            }
        };
        ffi::check(status);
        Factorized {
            handle,
            rows: a.nrows(),
            cols: a.ncols(),
            singular: if info < 0 { None } else { Some(info as usize) },
            _marker: PhantomData,
        }
    }
}

impl<A, S> Factorized<A, S>
where
    A: Scalar,
    S: Data<Elem = A>,
{
    fn view(&self, which: i32, rows: usize, cols: usize) -> Array2<A> {
        let mut out = Array2::<A>::zeros((rows, cols));
        ffi::check(unsafe { ffi::lair_b200_lu_view(self.handle, which, out.as_mut_ptr().cast(), cols as i64, 1) });
        out
    }

    /// lu.rs:28-39
    pub fn p(&self) -> Array2<A> {
        self.view(VIEW_P, self.rows, self.rows)
    }
    /// lu.rs:42-57
    pub fn l(&self) -> Array2<A> {
        self.view(VIEW_L, self.rows, self.rows.min(self.cols))
    }
    /// lu.rs:60-72
    pub fn u(&self) -> Array2<A> {
        self.view(VIEW_U, self.rows.min(self.cols), self.cols)
    }
    /// lu.rs:75-77
    pub fn is_singular(&self) -> bool {
        self.singular.is_some()
    }

    /// lu.rs:87-98: only `b` travels to the device and `x` back.
    pub fn solve<SB>(&self, b: &ArrayBase<SB, Ix1>) -> Result<Array1<A>, InvalidInput>
    where
        SB: Data<Elem = A>,
    {
        if b.len() != self.rows {
            return Err(InvalidInput::Shape(format!("b must have {} elements", self.rows)));
        }
        let mut x = Array1::<A>::zeros(self.rows);
        ffi::check(unsafe {
            ffi::lair_b200_lu_solve(self.handle, 1, b.as_ptr().cast(), b.stride_of(Axis(0)) as i64, 1,
                                    x.as_mut_ptr().cast(), 1, 1)
        });
        Ok(x)
    }

    /// lu.rs:107-153 (square / tall inputs: every column of the result is P*L).
    pub fn into_pl(self) -> Array2<A> {
        self.view(VIEW_PL, self.rows, self.rows.min(self.cols))
    }
}
