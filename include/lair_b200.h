/*
 * lair_b200.h -- C ABI of the B200-native LU path (getrf / getrs) behind lair's API.
 *
 * The reference (vinesystems/lair v0.8.0, pure Rust) has NO FFI for this path; the
 * drop-in boundary is two generic Rust signatures and their callers:
 *
 *   lapack::getrf<A>(a: ArrayViewMut2<A>) -> (Vec<usize>, Option<usize>)   src/lapack/getrf.rs:12-27
 *   lapack::getrs<A,SA,SB>(a, p: &[usize], b: &ArrayBase<SB,Ix1>) -> Array1<A>   src/lapack/getrs.rs:12-38
 *   lapack::laswp<T>(ncols, a, row_stride, col_stride, begin, piv)          src/lapack/laswp.rs:11-40
 *   lu::Factorized::from / ::solve / ::is_singular                          src/decomposition/lu.rs:156-171, 87-98, 75-77
 *   equation::solve                                                          src/equation.rs:32-60
 *
 * Every entry point below is what a thin Rust shim (rust/src/ffi.rs)
 * binds underneath those unchanged signatures.  Plain pointers and sizes only; no
 * torch / C++ types.  All functions return a lair_b200_status (0 = ok); the message of
 * the last failure on the calling thread is available from lair_b200_last_error().
 * There is NO CPU fallback: without a usable sm_100 device every compute entry point
 * returns LAIR_B200_ERR_NO_DEVICE (the Rust shim turns non-zero into panic!, because the
 * reference signatures have no error channel for runtime failure).
 *
 * Conventions shared with the reference:
 *   - strides are in ELEMENTS and may be negative or transposed (getrf.rs:52-54, tests :430-483);
 *   - ipiv is the 0-based LAPACK-style sequential interchange vector of length min(m,n)
 *     (getrf.rs:18-19): row i was swapped with row ipiv[i] at step i;
 *   - *info = -1 for None, else the LAST step whose pivot was exactly zero
 *     (getrf.rs:72-73, 168-169; differs from LAPACK's first-zero, 1-based info);
 *   - pivot search = first index of max |re|+|im|, NaN never wins (src/blas/iamax.rs:6-21);
 *   - complex scalars are interleaved (re, im) pairs, as num_complex::Complex<T> / C99.
 */
#ifndef LAIR_B200_H
#define LAIR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define LAIR_B200_API __attribute__((visibility("default")))
#else
#define LAIR_B200_API
#endif

typedef enum lair_b200_status {
    LAIR_B200_OK = 0,
    LAIR_B200_ERR_INVALID = 1,     /* bad argument (shape / stride / null pointer)            */
    LAIR_B200_ERR_NO_DEVICE = 2,   /* no CUDA device, or device is not sm_100                 */
    LAIR_B200_ERR_CUDA = 3,        /* a CUDA runtime call or kernel failed                    */
    LAIR_B200_ERR_ALLOC = 4,       /* host or device allocation failed                        */
    LAIR_B200_ERR_UNSUPPORTED = 5, /* shape/type combination not implemented on the device    */
    LAIR_B200_ERR_NCCL = 6         /* NCCL could not be loaded or a collective failed         */
} lair_b200_status;

/* ---- context ------------------------------------------------------------------------ */
/* Library version: major*10000 + minor*100 + patch. */
LAIR_B200_API int lair_b200_version(void);
/* Message for the last non-zero status returned on this thread ("" if none). */
LAIR_B200_API const char* lair_b200_last_error(void);
/* Number of visible CUDA devices (0 and LAIR_B200_ERR_NO_DEVICE when none). */
LAIR_B200_API int lair_b200_device_count(int* count);
/* Bind the process-wide context to `device` (creates streams and workspaces lazily;
 * idempotent for the same device).  Compute entry points call this with the current
 * device if it was never called. */
LAIR_B200_API int lair_b200_init(int device);
LAIR_B200_API int lair_b200_shutdown(void);

/* ---- host-pointer drop-ins (what the Rust shim binds) ---------------------------------
 * getrf: factor the m x n host matrix `a` in place through its strides; H2D / D2H copies
 * are inside the call.  Replaces src/lapack/getrf.rs:12-27 (and through it
 * lu::Factorized::from, src/decomposition/lu.rs:163-170).
 *   ipiv: min(m,n) int64 (widen to usize in the shim); info: see above.             */
LAIR_B200_API int lair_b200_sgetrf(int64_t m, int64_t n, float* a, int64_t row_stride, int64_t col_stride, int64_t* ipiv, int64_t* info);
LAIR_B200_API int lair_b200_dgetrf(int64_t m, int64_t n, double* a, int64_t row_stride, int64_t col_stride, int64_t* ipiv, int64_t* info);
LAIR_B200_API int lair_b200_cgetrf(int64_t m, int64_t n, void* a, int64_t row_stride, int64_t col_stride, int64_t* ipiv, int64_t* info);
LAIR_B200_API int lair_b200_zgetrf(int64_t m, int64_t n, void* a, int64_t row_stride, int64_t col_stride, int64_t* ipiv, int64_t* info);

/* getrs: X = U^-1 L^-1 P B for nrhs right-hand sides.  Replaces src/lapack/getrs.rs:12-38
 * (nrhs = 1, b_cs / x_cs ignored) and adds the multi-RHS form the reference lacks
 * (SURVEY 0.5): column r of X equals the reference's getrs on column r of B.
 * `lu` is n x n (any strides), ipiv has n entries, b / x are n x nrhs (any strides; x may
 * alias b only if their strides are identical).  ipiv must be getrf's sequential
 * interchanges, i <= ipiv[i] < n (the reference's laswp accepts any in-range entry; the
 * device kernels track only rows at or below the current step, so anything else is refused
 * with LAIR_B200_ERR_INVALID instead of being mis-applied).                           */
LAIR_B200_API int lair_b200_sgetrs(int64_t n, int64_t nrhs, const float* lu, int64_t lu_rs, int64_t lu_cs, const int64_t* ipiv, const float* b, int64_t b_rs, int64_t b_cs, float* x, int64_t x_rs, int64_t x_cs);
LAIR_B200_API int lair_b200_dgetrs(int64_t n, int64_t nrhs, const double* lu, int64_t lu_rs, int64_t lu_cs, const int64_t* ipiv, const double* b, int64_t b_rs, int64_t b_cs, double* x, int64_t x_rs, int64_t x_cs);
LAIR_B200_API int lair_b200_cgetrs(int64_t n, int64_t nrhs, const void* lu, int64_t lu_rs, int64_t lu_cs, const int64_t* ipiv, const void* b, int64_t b_rs, int64_t b_cs, void* x, int64_t x_rs, int64_t x_cs);
LAIR_B200_API int lair_b200_zgetrs(int64_t n, int64_t nrhs, const void* lu, int64_t lu_rs, int64_t lu_cs, const int64_t* ipiv, const void* b, int64_t b_rs, int64_t b_cs, void* x, int64_t x_rs, int64_t x_cs);

/* gesv: the whole of equation::solve (src/equation.rs:32-60) in one call -- A is copied
 * (never modified), factored on the device, and the factors stay device-resident for the
 * solve (no D2H/H2D of L\U in between).  *info as getrf; when *info >= 0 the reference
 * returns Err(InvalidInput::Value) and x is left untouched.                          */
LAIR_B200_API int lair_b200_sgesv(int64_t n, int64_t nrhs, const float* a, int64_t a_rs, int64_t a_cs, const float* b, int64_t b_rs, int64_t b_cs, float* x, int64_t x_rs, int64_t x_cs, int64_t* info);
LAIR_B200_API int lair_b200_dgesv(int64_t n, int64_t nrhs, const double* a, int64_t a_rs, int64_t a_cs, const double* b, int64_t b_rs, int64_t b_cs, double* x, int64_t x_rs, int64_t x_cs, int64_t* info);

/* ---- device-resident factors: lu::Factorized (src/decomposition/lu.rs:12-20) ------------------
 * The reference's Factorized owns L\U, the pivots and the singular flag between `from` and `solve`
 * (lu.rs:156-171, 87-98); its getrs signature forces the host-pointer entry points above to upload
 * L\U again for every right-hand side.  A handle keeps them in HBM instead:
 *   lu_factor   Factorized::from: A (host, any strides, NOT modified -- the reference consumes its
 *               argument, so the caller cannot observe it) is uploaded and factored; *info as getrf.
 *   lu_solve    Factorized::solve for nrhs right-hand sides (b: n x nrhs through its strides; the
 *               reference's Ix1 b is nrhs = 1).  Needs a square factorization (getrs.rs:18-20);
 *               a singular one is solved as the reference's getrs would (division by the zero pivot).
 *   lu_pivots   the interchange vector (int64[min(m,n)], 0-based, sequential).
 *   lu_factors  packed L\U to the host (m x n through the strides) -- the `lu` field.
 *   lu_view     Factorized::l / u / p / into_pl (lu.rs:42-57, 60-72, 28-39, 107-153) built by
 *               device kernels from the resident factors; out is m x k, k x n, m x m, m x k.
 *   lu_shape    m, n, dtype code (0 f32, 1 f64, 2 c32, 3 c64), info.
 *   lu_destroy  frees the device buffers.  Handles are not thread-safe individually; calls
 *               serialise on the library stream like the other host-pointer entry points.   */
typedef struct lair_b200_lu* lair_b200_lu_t;
enum { LAIR_LU_VIEW_L = 0, LAIR_LU_VIEW_U = 1, LAIR_LU_VIEW_P = 2, LAIR_LU_VIEW_PL = 3 };
LAIR_B200_API int lair_b200_slu_factor(int64_t m, int64_t n, const float* a, int64_t rs, int64_t cs, lair_b200_lu_t* handle, int64_t* info);
LAIR_B200_API int lair_b200_dlu_factor(int64_t m, int64_t n, const double* a, int64_t rs, int64_t cs, lair_b200_lu_t* handle, int64_t* info);
LAIR_B200_API int lair_b200_clu_factor(int64_t m, int64_t n, const void* a, int64_t rs, int64_t cs, lair_b200_lu_t* handle, int64_t* info);
LAIR_B200_API int lair_b200_zlu_factor(int64_t m, int64_t n, const void* a, int64_t rs, int64_t cs, lair_b200_lu_t* handle, int64_t* info);
LAIR_B200_API int lair_b200_lu_solve(lair_b200_lu_t handle, int64_t nrhs, const void* b, int64_t b_rs, int64_t b_cs, void* x, int64_t x_rs, int64_t x_cs);
LAIR_B200_API int lair_b200_lu_pivots(lair_b200_lu_t handle, int64_t* ipiv);
LAIR_B200_API int lair_b200_lu_factors(lair_b200_lu_t handle, void* lu, int64_t rs, int64_t cs);
LAIR_B200_API int lair_b200_lu_view(lair_b200_lu_t handle, int view, void* out, int64_t rs, int64_t cs);
LAIR_B200_API int lair_b200_lu_shape(lair_b200_lu_t handle, int64_t* m, int64_t* n, int* dtype, int64_t* info);
LAIR_B200_API int lair_b200_lu_destroy(lair_b200_lu_t handle);

/* Batched LU of `batch` independent, contiguous, row-major n x n matrices (n <= 32),
 * i.e. `batch` calls of getrf.rs:12-27 on standard-layout inputs; results are bit-identical
 * to the reference's row-major body (getrf.rs:46-120).  ipiv: batch*n int32, info: batch
 * int32 (-1 = None).                                                                   */
LAIR_B200_API int lair_b200_sgetrf_batched(int64_t batch, int64_t n, float* a, int32_t* ipiv, int32_t* info);
LAIR_B200_API int lair_b200_dgetrf_batched(int64_t batch, int64_t n, double* a, int32_t* ipiv, int32_t* info);
/* The same call spread over the first `ngpu` devices of the node from ONE process (independent units,
 * contiguous slices of the batch, no collective): device d factors matrices [d*batch/ngpu, ...) with its own
 * H2D / factor / D2H pipeline over its own PCIe link.  ngpu = 1 is the call above on device 0's streams.  Results are
 * bit-identical to the single-GPU entry (same kernels).  Pinned host memory lets the devices overlap; pageable memory
 * still works (the driver stages the copies).  LAIR_B200_ERR_INVALID when more GPUs are requested than are visible. */
LAIR_B200_API int lair_b200_sgetrf_batched_mg(int64_t batch, int64_t n, float* a, int32_t* ipiv, int32_t* info, int ngpu);
LAIR_B200_API int lair_b200_dgetrf_batched_mg(int64_t batch, int64_t n, double* a, int32_t* ipiv, int32_t* info, int ngpu);

/* ---- device-resident variants (device pointers, caller's stream, no copies) -----------
 * Used by the benchmark for the HBM-resident number and by callers that keep data on the
 * GPU.  Device layout is ROW-MAJOR with leading dimension `lda` (elements, >= n).
 * d_ipiv: int32[min(m,n)] global 0-based rows; d_info: int32[1].  `stream` is a
 * cudaStream_t passed as void* (NULL = default stream).  Asynchronous w.r.t. the host.
 * The *_dev entry points share the context's workspaces (panel exchange slots, lookahead
 * stream and events, scratch): use them from ONE stream at a time and not concurrently with a
 * host-pointer entry point -- order calls on different streams with events.
 * *laswp_dev takes the same kind of pivots as getrs (i <= d_ipiv[i]).                  */
LAIR_B200_API int lair_b200_sgetrf_dev(int64_t m, int64_t n, float* d_a, int64_t lda, int32_t* d_ipiv, int32_t* d_info, void* stream);
LAIR_B200_API int lair_b200_dgetrf_dev(int64_t m, int64_t n, double* d_a, int64_t lda, int32_t* d_ipiv, int32_t* d_info, void* stream);
/* Complex<f32> / Complex<f64> (interleaved re, im; src/scalar.rs:370-386): beyond 128 x 128 a blocked sweep whose
 * trailing update is one real GEMM on packed operands (blocked_cx.cu).                                            */
LAIR_B200_API int lair_b200_cgetrf_dev(int64_t m, int64_t n, void* d_a, int64_t lda, int32_t* d_ipiv, int32_t* d_info, void* stream);
LAIR_B200_API int lair_b200_zgetrf_dev(int64_t m, int64_t n, void* d_a, int64_t lda, int32_t* d_ipiv, int32_t* d_info, void* stream);
/* In-place solve on the device: d_b (n x nrhs, row-major, ldb) is overwritten with X. */
LAIR_B200_API int lair_b200_sgetrs_dev(int64_t n, int64_t nrhs, const float* d_lu, int64_t lda, const int32_t* d_ipiv, float* d_b, int64_t ldb, void* stream);
LAIR_B200_API int lair_b200_dgetrs_dev(int64_t n, int64_t nrhs, const double* d_lu, int64_t lda, const int32_t* d_ipiv, double* d_b, int64_t ldb, void* stream);
LAIR_B200_API int lair_b200_sgetrf_batched_dev(int64_t batch, int64_t n, float* d_a, int32_t* d_ipiv, int32_t* d_info, void* stream);
LAIR_B200_API int lair_b200_dgetrf_batched_dev(int64_t batch, int64_t n, double* d_a, int32_t* d_ipiv, int32_t* d_info, void* stream);

/* Building blocks of the blocked factorization, exported for parity tests and profiling
 * (each mirrors one reference routine; all device-resident, row-major):
 *   laswp  src/lapack/laswp.rs:11-40   rows of d_a[:, 0..ncols) swapped by d_ipiv[k0..k1)
 *          (entries are global row indices relative to d_a's row 0)
 *   trsm   src/blas/trsm.rs:6-22       B <- L^-1 B, L = unit-lower k x k at d_l
 *   gemm   src/blas/gemm.rs:6-32       C -= A * B  (alpha = -1, no conjugation: the LU call site
 *                                       src/lapack/getrf.rs:289-296)                     */
LAIR_B200_API int lair_b200_dlaswp_dev(int64_t ncols, double* d_a, int64_t lda, int64_t k0, int64_t k1, const int32_t* d_ipiv, void* stream);
LAIR_B200_API int lair_b200_slaswp_dev(int64_t ncols, float* d_a, int64_t lda, int64_t k0, int64_t k1, const int32_t* d_ipiv, void* stream);
LAIR_B200_API int lair_b200_dtrsm_dev(int64_t k, int64_t ncols, const double* d_l, int64_t ldl, double* d_b, int64_t ldb, void* stream);
LAIR_B200_API int lair_b200_strsm_dev(int64_t k, int64_t ncols, const float* d_l, int64_t ldl, float* d_b, int64_t ldb, void* stream);
LAIR_B200_API int lair_b200_dgemm_minus_dev(int64_t m, int64_t n, int64_t k, const double* d_a, int64_t lda, const double* d_b, int64_t ldb, double* d_c, int64_t ldc, void* stream);
LAIR_B200_API int lair_b200_sgemm_minus_dev(int64_t m, int64_t n, int64_t k, const float* d_a, int64_t lda, const float* d_b, int64_t ldb, float* d_c, int64_t ldc, void* stream);

/* ---- multi-GPU: one process per GPU, NCCL over NVLink --------------------------
 * One large LU on a 1-D block-cyclic COLUMN distribution: global block column j (width nb)
 * lives on rank j mod P as local block j / P; d_a_local is n rows x local_cols, row-major with
 * leading dimension lda.  Per block column the owner factors the panel and broadcasts the packed
 * panel + nb pivots (ncclBroadcast); every rank updates its own columns; one block of lookahead.
 * d_ipiv (n int32, replicated on every rank) and d_info as in getrf_dev.
 * Communicator set-up: rank 0 calls mg_unique_id (128 bytes), the host layer distributes the
 * bytes (torch.distributed or any other channel), every rank calls mg_init. */
LAIR_B200_API int lair_b200_mg_unique_id(void* id128);
LAIR_B200_API int lair_b200_mg_init(int rank, int nranks, const void* id128);
LAIR_B200_API int lair_b200_mg_finalize(void);
/* Measurement aid: with enable != 0 the next *getrf_mg_dev calls record 8 CUDA timing events per block step on this rank
 * (main stream: panel landed / own next block updated / trailing update done / left interchanges done; panel +
 * communication stream: start / factored / packed / broadcast done).  mg_timeline_read synchronises the device and
 * returns out[b * 8 + p] = milliseconds since the start of the last call (NaN: point not taken on this rank). */
LAIR_B200_API int lair_b200_mg_timeline(int enable);
LAIR_B200_API int lair_b200_mg_timeline_read(float* out, int64_t cap, int64_t* nblk);
LAIR_B200_API int lair_b200_dgetrf_mg_dev(int64_t n, int64_t nb, double* d_a_local, int64_t lda, int32_t* d_ipiv, int32_t* d_info, void* stream);
LAIR_B200_API int lair_b200_sgetrf_mg_dev(int64_t n, int64_t nb, float* d_a_local, int64_t lda, int32_t* d_ipiv, int32_t* d_info, void* stream);

/* Tuning knobs (also read from the environment at init: LAIR_B200_NB, LAIR_B200_SMALL_N).
 * Behaviour: "nb" (outer block width, 0 = by remaining size), "nb_t1"/"nb_t2" (its thresholds),
 * "small_n", "lookahead", "stream_cols"/"stream_join_div" (chunked upload of the host-pointer
 * entry points), "batched_chunk" (matrices per pipelined chunk of the host-pointer batched LU).  Kernel variants kept for the parity tests and A/B measurements:
 * "batched_cfg", "panel_cluster", "panel_rpt", "panel_group", "panel_exchange", "panel_w64",
 * "panel_timing", "gemm_cfg", "fuse_swap_trsm", "trsm_dataflow", "trsm_rb", "laswp_perm"
 * (meanings next to Options in csrc/common.cuh).  Every variant computes the same result to the
 * parity bars of DESIGN.md section 5.  Unknown names return LAIR_B200_ERR_INVALID. */
LAIR_B200_API int lair_b200_set_option(const char* name, int64_t value);
LAIR_B200_API int lair_b200_get_option(const char* name, int64_t* value);
/* Number of kernels this library has launched since init (for gpu_launches accounting). */
LAIR_B200_API int64_t lair_b200_launch_count(void);

/* Device-resident (_dev) calls are asynchronous.  The kernels that wait on other CTAs through
 * global memory (panels taller than one cluster, the dataflow triangular solves of getrs) bound
 * their waits so a lost peer cannot hang the GPU; a wait that runs out raises a device-side fault
 * and the results are invalid.  check_fault waits for `stream`, and returns LAIR_B200_ERR_CUDA (with
 * a message) if a fault was raised since the last check, LAIR_B200_OK otherwise.  The host-pointer
 * entry points perform this check themselves before they return. */
LAIR_B200_API int lair_b200_check_fault(void* stream);
/* Optional live timing of kernel families with CUDA events on the launching stream:
 * begin() arms it, end() synchronises and accumulates, get() returns for one family
 * ("gemm", "panel", "laswp", "trsm", "batched", "small") the summed device time (ms), the
 * number of launches and their algorithmic work (flops for gemm/trsm/small, bytes otherwise). */
LAIR_B200_API int lair_b200_profile_begin(void);
LAIR_B200_API int lair_b200_profile_end(void);
LAIR_B200_API int lair_b200_profile_get(const char* family, double* ms, int64_t* launches, double* work);
/* Debug aid: with option "panel_timing" = 1 the cluster panel kernel accumulates SM-cycle counts
 * of its per-column phases; out8 = {candidate, block barrier, push, cluster barrier, winner,
 * update, columns, 0}. */
LAIR_B200_API int lair_b200_debug_panel_timing(long long* out8, int clear);

/* ---- Householder QR -------------------------------------------------------------
 * geqrf: lapack::geqrf (src/lapack/geqrf.rs:9-30) -- `a` (m x n, element strides rs / cs, any layout) is overwritten
 * with R on and above the diagonal and the reflector vectors below it; tau receives min(m, n) entries.
 * qr_q: qr::Factorized::q (src/decomposition/qr.rs:27-59) -- the m x m unitary factor from the factored matrix and
 * tau, written to `q` (element strides q_rs / q_cs).  R is the upper triangle of the factored matrix (qr.rs:62-70).
 * Complex types: interleaved (re, im).  geqrf_dev: device pointers, row-major with leading dimension lda.          */
LAIR_B200_API int lair_b200_sgeqrf(int64_t m, int64_t n, float* a, int64_t rs, int64_t cs, float* tau);
LAIR_B200_API int lair_b200_dgeqrf(int64_t m, int64_t n, double* a, int64_t rs, int64_t cs, double* tau);
LAIR_B200_API int lair_b200_cgeqrf(int64_t m, int64_t n, void* a, int64_t rs, int64_t cs, void* tau);
LAIR_B200_API int lair_b200_zgeqrf(int64_t m, int64_t n, void* a, int64_t rs, int64_t cs, void* tau);
LAIR_B200_API int lair_b200_sqr_q(int64_t m, int64_t n, const float* qr, int64_t rs, int64_t cs, const float* tau, float* q, int64_t q_rs, int64_t q_cs);
LAIR_B200_API int lair_b200_dqr_q(int64_t m, int64_t n, const double* qr, int64_t rs, int64_t cs, const double* tau, double* q, int64_t q_rs, int64_t q_cs);
LAIR_B200_API int lair_b200_cqr_q(int64_t m, int64_t n, const void* qr, int64_t rs, int64_t cs, const void* tau, void* q, int64_t q_rs, int64_t q_cs);
LAIR_B200_API int lair_b200_zqr_q(int64_t m, int64_t n, const void* qr, int64_t rs, int64_t cs, const void* tau, void* q, int64_t q_rs, int64_t q_cs);
LAIR_B200_API int lair_b200_sgeqrf_dev(int64_t m, int64_t n, float* d_a, int64_t lda, float* d_tau, void* stream);
LAIR_B200_API int lair_b200_dgeqrf_dev(int64_t m, int64_t n, double* d_a, int64_t lda, double* d_tau, void* stream);
LAIR_B200_API int lair_b200_cgeqrf_dev(int64_t m, int64_t n, void* d_a, int64_t lda, void* d_tau, void* stream);
LAIR_B200_API int lair_b200_zgeqrf_dev(int64_t m, int64_t n, void* d_a, int64_t lda, void* d_tau, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LAIR_B200_H */
