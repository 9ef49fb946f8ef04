//! lair's LU path on NVIDIA B200: `decomposition::lu`, `equation::solve` and the crate-private
//! `lapack::{getrf, getrs, laswp}` keep the reference's signatures; the arithmetic runs in
//! hand-written CUDA (sm_100a) behind the C ABI declared in include/lair_b200.h.
//!
//! `decomposition/lu.rs`, `equation.rs` and `scalar.rs` are the reference's files unchanged
//! (they only call `lapack::*`), so they are not duplicated in this repository: drop the three
//! files of `src/lapack/` and `ffi.rs` into the reference tree, add `build.rs`, done.
//! Optional, same public surface: `equation.rs` (one `gesv` call for f32 / f64 instead of copy +
//! factor + solve), `decomposition/lu_device.rs` (factors stay in HBM behind a handle) and
//! `lapack/getrf_batched.rs` (the loop over small matrices as one call) replace / join the
//! reference's files of the same module paths.
//!
//! Module layout when dropped into the reference tree (src/lib.rs:1-30 of the reference declares the same names):
//! `pub mod decomposition { pub mod lu; ... }`, `pub mod equation;`, `mod scalar;`, `mod lapack;` stay the reference's;
//! this crate adds `mod ffi;`.  The two optional replacements are declared below so that `equation.rs`'s
//! `use crate::decomposition::lu` resolves in this source-only tree as well.
mod ffi;
mod lapack;

pub mod decomposition {
    /// device-resident `Factorized<A, S>` (replaces the reference's src/decomposition/lu.rs when factors should stay in HBM)
    #[path = "lu_device.rs"]
    pub mod lu;
}
pub mod equation;
// `Scalar`, `Real` and `InvalidInput` are the reference's own (src/scalar.rs, src/lib.rs): not duplicated here.

#[cfg(feature = "bench-lapack")]
#[doc(hidden)]
pub mod bench_lapack {
    pub use crate::lapack::*;
}
