//! Raw bindings of include/lair_b200.h (the C ABI of the CUDA library).
#![allow(non_camel_case_types)]
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

extern "C" {
    pub fn lair_b200_last_error() -> *const c_char;
    pub fn lair_b200_init(device: c_int) -> c_int;
    pub fn lair_b200_shutdown() -> c_int;

    pub fn lair_b200_sgetrf(m: i64, n: i64, a: *mut f32, rs: i64, cs: i64, ipiv: *mut i64, info: *mut i64) -> c_int;
    pub fn lair_b200_dgetrf(m: i64, n: i64, a: *mut f64, rs: i64, cs: i64, ipiv: *mut i64, info: *mut i64) -> c_int;
    pub fn lair_b200_cgetrf(m: i64, n: i64, a: *mut c_void, rs: i64, cs: i64, ipiv: *mut i64, info: *mut i64) -> c_int;
    pub fn lair_b200_zgetrf(m: i64, n: i64, a: *mut c_void, rs: i64, cs: i64, ipiv: *mut i64, info: *mut i64) -> c_int;

    pub fn lair_b200_sgetrs(n: i64, nrhs: i64, lu: *const f32, lu_rs: i64, lu_cs: i64, ipiv: *const i64,
                            b: *const f32, b_rs: i64, b_cs: i64, x: *mut f32, x_rs: i64, x_cs: i64) -> c_int;
    pub fn lair_b200_dgetrs(n: i64, nrhs: i64, lu: *const f64, lu_rs: i64, lu_cs: i64, ipiv: *const i64,
                            b: *const f64, b_rs: i64, b_cs: i64, x: *mut f64, x_rs: i64, x_cs: i64) -> c_int;
    pub fn lair_b200_cgetrs(n: i64, nrhs: i64, lu: *const c_void, lu_rs: i64, lu_cs: i64, ipiv: *const i64,
                            b: *const c_void, b_rs: i64, b_cs: i64, x: *mut c_void, x_rs: i64, x_cs: i64) -> c_int;
    pub fn lair_b200_zgetrs(n: i64, nrhs: i64, lu: *const c_void, lu_rs: i64, lu_cs: i64, ipiv: *const i64,
                            b: *const c_void, b_rs: i64, b_cs: i64, x: *mut c_void, x_rs: i64, x_cs: i64) -> c_int;

    // device-resident factors (lu::Factorized behind a handle; include/lair_b200.h)
    pub fn lair_b200_slu_factor(m: i64, n: i64, a: *const f32, rs: i64, cs: i64, handle: *mut *mut c_void, info: *mut i64) -> c_int;
    pub fn lair_b200_dlu_factor(m: i64, n: i64, a: *const f64, rs: i64, cs: i64, handle: *mut *mut c_void, info: *mut i64) -> c_int;
    pub fn lair_b200_clu_factor(m: i64, n: i64, a: *const c_void, rs: i64, cs: i64, handle: *mut *mut c_void, info: *mut i64) -> c_int;
    pub fn lair_b200_zlu_factor(m: i64, n: i64, a: *const c_void, rs: i64, cs: i64, handle: *mut *mut c_void, info: *mut i64) -> c_int;
    pub fn lair_b200_lu_solve(handle: *mut c_void, nrhs: i64, b: *const c_void, b_rs: i64, b_cs: i64,
                              x: *mut c_void, x_rs: i64, x_cs: i64) -> c_int;
    pub fn lair_b200_lu_pivots(handle: *mut c_void, ipiv: *mut i64) -> c_int;
    pub fn lair_b200_lu_factors(handle: *mut c_void, lu: *mut c_void, rs: i64, cs: i64) -> c_int;
    pub fn lair_b200_lu_view(handle: *mut c_void, view: c_int, out: *mut c_void, rs: i64, cs: i64) -> c_int;
    pub fn lair_b200_lu_destroy(handle: *mut c_void) -> c_int;

    // Householder QR (include/lair_b200.h, "Householder QR"): geqrf in place + tau, q from the factors
    pub fn lair_b200_sgeqrf(m: i64, n: i64, a: *mut f32, rs: i64, cs: i64, tau: *mut f32) -> c_int;
    pub fn lair_b200_dgeqrf(m: i64, n: i64, a: *mut f64, rs: i64, cs: i64, tau: *mut f64) -> c_int;
    pub fn lair_b200_cgeqrf(m: i64, n: i64, a: *mut c_void, rs: i64, cs: i64, tau: *mut c_void) -> c_int;
    pub fn lair_b200_zgeqrf(m: i64, n: i64, a: *mut c_void, rs: i64, cs: i64, tau: *mut c_void) -> c_int;
    pub fn lair_b200_sqr_q(m: i64, n: i64, qr: *const f32, rs: i64, cs: i64, tau: *const f32, q: *mut f32, q_rs: i64, q_cs: i64) -> c_int;
    pub fn lair_b200_dqr_q(m: i64, n: i64, qr: *const f64, rs: i64, cs: i64, tau: *const f64, q: *mut f64, q_rs: i64, q_cs: i64) -> c_int;
    pub fn lair_b200_cqr_q(m: i64, n: i64, qr: *const c_void, rs: i64, cs: i64, tau: *const c_void, q: *mut c_void, q_rs: i64, q_cs: i64) -> c_int;
    pub fn lair_b200_zqr_q(m: i64, n: i64, qr: *const c_void, rs: i64, cs: i64, tau: *const c_void, q: *mut c_void, q_rs: i64, q_cs: i64) -> c_int;

    // context queries (include/lair_b200.h, "context")
    pub fn lair_b200_version() -> c_int;
    pub fn lair_b200_device_count(count: *mut c_int) -> c_int;
    pub fn lair_b200_check_fault(stream: *mut c_void) -> c_int;

    // gesv: the whole of equation::solve (src/equation.rs:32-60) in one call; the factors never leave HBM
    pub fn lair_b200_sgesv(n: i64, nrhs: i64, a: *const f32, a_rs: i64, a_cs: i64, b: *const f32, b_rs: i64, b_cs: i64,
                           x: *mut f32, x_rs: i64, x_cs: i64, info: *mut i64) -> c_int;
    pub fn lair_b200_dgesv(n: i64, nrhs: i64, a: *const f64, a_rs: i64, a_cs: i64, b: *const f64, b_rs: i64, b_cs: i64,
                           x: *mut f64, x_rs: i64, x_cs: i64, info: *mut i64) -> c_int;

    // batched LU of contiguous row-major n x n matrices, n <= 32 (BASELINE configs[2]); int32 pivots / info per matrix
    pub fn lair_b200_sgetrf_batched(batch: i64, n: i64, a: *mut f32, ipiv: *mut i32, info: *mut i32) -> c_int;
    pub fn lair_b200_dgetrf_batched(batch: i64, n: i64, a: *mut f64, ipiv: *mut i32, info: *mut i32) -> c_int;
    // the same over the first `ngpu` devices of the node, one process (contiguous slices of the batch, no collective)
    pub fn lair_b200_sgetrf_batched_mg(batch: i64, n: i64, a: *mut f32, ipiv: *mut i32, info: *mut i32, ngpu: c_int) -> c_int;
    pub fn lair_b200_dgetrf_batched_mg(batch: i64, n: i64, a: *mut f64, ipiv: *mut i32, info: *mut i32, ngpu: c_int) -> c_int;

    // one large LU over several GPUs, one process per GPU (BASELINE configs[3]); device pointers
    pub fn lair_b200_mg_unique_id(id128: *mut c_void) -> c_int;
    pub fn lair_b200_mg_init(rank: c_int, nranks: c_int, id128: *const c_void) -> c_int;
    pub fn lair_b200_mg_finalize() -> c_int;
    pub fn lair_b200_mg_timeline(enable: c_int) -> c_int;
    pub fn lair_b200_mg_timeline_read(out: *mut f32, cap: i64, nblk: *mut i64) -> c_int;
    pub fn lair_b200_sgetrf_mg_dev(n: i64, nb: i64, d_a_local: *mut f32, lda: i64, d_ipiv: *mut i32, d_info: *mut i32,
                                   stream: *mut c_void) -> c_int;
    pub fn lair_b200_dgetrf_mg_dev(n: i64, nb: i64, d_a_local: *mut f64, lda: i64, d_ipiv: *mut i32, d_info: *mut i32,
                                   stream: *mut c_void) -> c_int;
}

/// The reference signatures have no error channel for runtime failure, so a non-zero status
/// (no device, CUDA error, allocation failure) becomes a panic carrying the library's message.
pub(crate) fn check(status: c_int) {
    if status != 0 {
        let msg = unsafe { CStr::from_ptr(lair_b200_last_error()) }.to_string_lossy().into_owned();
        panic!("lair_b200 status {status}: {msg}");
    }
}
