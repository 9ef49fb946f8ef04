//! Crate-private LAPACK-named drivers of the LU path and of geqrf (reference: src/lapack.rs:40-47).
mod geqrf;
mod getrf;
mod getrf_batched;
mod getrs;
mod laswp;

pub use geqrf::{geqrf, qr_q};
pub use getrf::getrf;
pub use getrf_batched::getrf_batched;
pub use getrs::getrs;
pub use laswp::laswp;
