//! Crate-private LAPACK-named drivers of the LU path (reference: src/lapack.rs:40-47).
mod getrf;
mod getrs;
mod laswp;

pub use getrf::getrf;
pub use getrs::getrs;
pub use laswp::laswp;
