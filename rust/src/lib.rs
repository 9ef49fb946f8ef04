//! lair's LU path on NVIDIA B200: `decomposition::lu`, `equation::solve` and the crate-private
//! `lapack::{getrf, getrs, laswp}` keep the reference's signatures; the arithmetic runs in
//! hand-written CUDA (sm_100a) behind the C ABI declared in include/lair_b200.h.
//!
//! `decomposition/lu.rs`, `equation.rs` and `scalar.rs` are the reference's files unchanged
//! (they only call `lapack::*`), so they are not duplicated in this repository: drop the three
//! files of `src/lapack/` and `ffi.rs` into the reference tree, add `build.rs`, done.
mod ffi;
mod lapack;

#[cfg(feature = "bench-lapack")]
#[doc(hidden)]
pub mod bench_lapack {
    pub use crate::lapack::*;
}
