// Shared declarations for the lair_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>

#include "../../include/lair_b200.h"

namespace lair {

// ---- error plumbing ---------------------------------------------------------------
void set_error(const char* fmt, ...);
extern thread_local int g_last_status;

#define LAIR_CUDA_CHECK(expr)                                                                  \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            ::lair::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return (_e == cudaErrorMemoryAllocation) ? LAIR_B200_ERR_ALLOC : LAIR_B200_ERR_CUDA; \
        }                                                                                      \
    } while (0)

#define LAIR_CHECK(expr)                 \
    do {                                 \
        int _s = (expr);                 \
        if (_s != LAIR_B200_OK) return _s; \
    } while (0)

#define LAIR_REQUIRE(cond, ...)              \
    do {                                     \
        if (!(cond)) {                       \
            ::lair::set_error(__VA_ARGS__);  \
            return LAIR_B200_ERR_INVALID;    \
        }                                    \
    } while (0)

// Counts kernels launched by this library (lair_b200_launch_count).
void count_launch(int n = 1);
#define LAIR_LAUNCH_CHECK()                      \
    do {                                         \
        ::lair::count_launch();                  \
        LAIR_CUDA_CHECK(cudaGetLastError());     \
    } while (0)

// ---- optional per-kernel-family timing (profile.cu) -----------------------------------------
enum ProfBucket { kProfGemm = 0, kProfPanel, kProfLaswp, kProfTrsm, kProfBatched, kProfSmall, kProfOther, kProfBuckets };
bool prof_enabled();
// Brackets the launches issued in its scope with CUDA events on `s`; `work` = algorithmic
// flops (gemm, trsm) or bytes (laswp, panel, batched) of those launches.
class ProfScope {
public:
    ProfScope(int bucket, cudaStream_t s, double work);
    ~ProfScope();
    ProfScope(const ProfScope&) = delete;
    ProfScope& operator=(const ProfScope&) = delete;

private:
    int bucket_;
    cudaStream_t stream_;
    double work_;
    cudaEvent_t e0_;
};

// ---- scalar layer: exact (never FMA-contracted) arithmetic ------------------------------
// The reference is Rust: `a -= l * u` is a rounded multiply followed by a rounded
// subtract (src/lapack/getrf.rs:86-87).  The *_rn intrinsics are never contracted by
// nvcc, so kernels that promise bit-identical results use these.
template <class T>
struct cx {
    T re, im;
};
using cxf = cx<float>;
using cxd = cx<double>;

template <class T> struct Ops;

template <> struct Ops<float> {
    using Real = float;
    static constexpr bool is_complex = false;
    __host__ __device__ static float zero() { return 0.f; }
    __host__ __device__ static float one() { return 1.f; }
    __device__ static float mul(float a, float b) { return __fmul_rn(a, b); }
    __device__ static float add(float a, float b) { return __fadd_rn(a, b); }
    __device__ static float sub(float a, float b) { return __fsub_rn(a, b); }
    __device__ static float div(float a, float b) { return __fdiv_rn(a, b); }
    __device__ static float recip(float a) { return __fdiv_rn(1.f, a); }
    __device__ static float abs1(float a) { return fabsf(a); }
    __device__ static bool is_zero(float a) { return a == 0.f; }
};
template <> struct Ops<double> {
    using Real = double;
    static constexpr bool is_complex = false;
    __host__ __device__ static double zero() { return 0.0; }
    __host__ __device__ static double one() { return 1.0; }
    __device__ static double mul(double a, double b) { return __dmul_rn(a, b); }
    __device__ static double add(double a, double b) { return __dadd_rn(a, b); }
    __device__ static double sub(double a, double b) { return __dsub_rn(a, b); }
    __device__ static double div(double a, double b) { return __ddiv_rn(a, b); }
    __device__ static double recip(double a) { return __ddiv_rn(1.0, a); }
    __device__ static double abs1(double a) { return fabs(a); }
    __device__ static bool is_zero(double a) { return a == 0.0; }
};
// Complex follows num-complex 0.4's textbook formulas (the reference's dependency,
// Cargo.toml:24), exactly like oracle/lair_oracle.hpp.
template <class R> struct Ops<cx<R>> {
    using Real = R;
    using O = Ops<R>;
    static constexpr bool is_complex = true;
    __host__ __device__ static cx<R> zero() { return {R(0), R(0)}; }
    __host__ __device__ static cx<R> one() { return {R(1), R(0)}; }
    __device__ static cx<R> mul(cx<R> a, cx<R> b) {
        return {O::sub(O::mul(a.re, b.re), O::mul(a.im, b.im)), O::add(O::mul(a.re, b.im), O::mul(a.im, b.re))};
    }
    __device__ static cx<R> add(cx<R> a, cx<R> b) { return {O::add(a.re, b.re), O::add(a.im, b.im)}; }
    __device__ static cx<R> sub(cx<R> a, cx<R> b) { return {O::sub(a.re, b.re), O::sub(a.im, b.im)}; }
    __device__ static cx<R> div(cx<R> a, cx<R> b) {
        R ns = O::add(O::mul(b.re, b.re), O::mul(b.im, b.im));
        R re = O::add(O::mul(a.re, b.re), O::mul(a.im, b.im));
        R im = O::sub(O::mul(a.im, b.re), O::mul(a.re, b.im));
        return {O::div(re, ns), O::div(im, ns)};
    }
    __device__ static cx<R> recip(cx<R> a) { return div(one(), a); }
    __device__ static R abs1(cx<R> a) { return O::add(O::abs1(a.re), O::abs1(a.im)); }
    __device__ static bool is_zero(cx<R> a) { return a.re == R(0) && a.im == R(0); }
};

// ---- process-wide context ----------------------------------------------------------------
struct Options {
    int64_t nb = 0;          // outer block width of the blocked factorization (0 = from the remaining size, below)
    int64_t nb_t1 = 0;       // nb = 0: blocks are 64 wide while <= nb_t1 columns remain (0 = 6144 for f64, 8192 for f32),
                             // 128 wide up to nb_t2, 256 beyond
    int64_t nb_t2 = 10240;   // thresholds measured on B200 (profiles/r1_bench_history.md)
    int64_t small_n = 128;   // max(m, n) handled by the single-CTA exact kernel
    int64_t lookahead = 1;   // overlap panel k+1 with trailing update k
    int64_t chain_on_p = 3072;  // lookahead: while more than this many columns remain, the next block's update runs on the
                             // panel stream (no cross-stream hand-over on the chain); 0 = always on the main stream
    int64_t batched_cfg = -1; // tuning variant of the batched kernel, -1 = measured best per type (batched_lu.cu)
    int64_t panel_cluster = 3;  // panels that fit one cluster: 3 record-carries-the-window kernel (panel_push.cu), 2 blocked DSMEM kernel with row pull, 1 row-per-thread DSMEM kernel, 0 global-memory exchange
    int64_t panel_rpt = 0;      // rows per thread of the cluster panel kernels (1, 2, 4); 0 = the measured best per shape (third generation: 2)
    int64_t panel_group = 4;    // columns per compiled group body of the cluster panel kernel (2, 4, 8)
    int64_t panel_timing = 0;   // debug: accumulate per-phase cycle counts in the cluster panel kernel
    int64_t panel_w64 = 1;      // panel_blocked: take a whole 64-column block in one launch when its rows fit
    int64_t panel_exchange = 1; // panel_blocked: 1 = st.async record push + winner-row pull, 0 = cluster barrier + pull
    int64_t stream_cols = 1024; // host-pointer getrf/gesv: upload the matrix in column chunks of this width and start
                                // factoring when the first has landed (0 = upload everything first)
    int64_t batched_chunk = 8192; // host-pointer batched LU: matrices per pipelined H2D / factor / D2H chunk (sweep: profiles/r1e_quick_batched_e2e.jsonl)
    int64_t stream_join_div = 4; // a chunk starting at column cs joins the sweep once cs / stream_join_div columns are factored
    int64_t laswp_perm = 1;    // getrs: apply P to the right-hand sides as one collapsed permutation (laswp_perm.cu)
    int64_t fuse_swap_trsm = 1; // block steps of width <= 128: one fused laswp+trsm launch (laswp_trsm.cu)
    int64_t trsm_dataflow = 2;  // f64 getrs: >= 1 flag-in-data dataflow solves with pre-inverted diagonal blocks (trsm_ll.cu),
                                // 0 recursive TRSM + GEMM (the flag-word first generation, 4.1 ms against 2.1 ms at n = 8192 / 64 RHS, was removed in round 2)
    int64_t trsm_rb = 32;       // row-block height of the flag-word dataflow solves (32 or 64; same speed, measured)
    int64_t cx_blocked = 1;     // complex beyond small_n: 1 blocked sweep (blocked_cx.cu), 2 the same with single-CTA leaf panels, 0 the single-CTA in-place kernel
    int64_t qr_blocked = 1;     // f32 / f64 geqrf with min(m, n) >= 64: 1 compact-WY blocks (qr_blocked.cu), 0 one reflector at a time
    int64_t gemm_cfg = 0;       // f64 GEMM tile: 0 auto, 1 big 128x64, 2 skinny 64x32, 3 128x128 (gemm_f64.cu)
    int64_t pair_k512 = 16384;  // f64 sweep: while more than this many columns remain, two 256-wide block steps share one K = 512 trailing GEMM (0 = never)
    int64_t pair_small = 0;     // f64 sweep: 128- / 64-wide block steps share one trailing GEMM while more than this many columns remain (0 = never; measured slower at n = 4096 ... 16384 for every threshold: the deferral costs more overlap than the deeper K buys, profiles/r2z_probe_pair_small.jsonl)
    int64_t pair_small_f32 = 0; // the same for f32
    int64_t trsm_strip = 2;     // trailing update on wide column ranges: laswp + one register-tiled triangle launch (trsm_strip.cu); 1 = only for k > 64, 0 = chain of fused 64-row launches
    int64_t drain_rows = 1;     // host-pointer getrf: finished rows go back to a pinned host array while the sweep runs (0 = one copy at the end)
    int64_t mg_signal_comm = 1; // multi-GPU LU: pivots travel first on a one-CTA communicator, so the panel's wide broadcast never waits on the device (mg.cu)
    int64_t sgemm_tf32 = 1;     // f32 GEMM: 1 = tcgen05 3xTF32 tensor-core path for large updates (gemm_tf32.cu), 0 = FP32 FMA kernel always
    int64_t gemm_raster = 8;    // f64 GEMM: tile columns per strip of the CTA order (1 = walk down M one tile column at a time)
};

struct Context {
    bool ready = false;
    int device = -1;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;        // library-owned stream for the host-pointer entry points
    cudaStream_t aux_stream = nullptr;    // second stream (lookahead / copy overlap)
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    int* d_fault = nullptr;               // device word raised by a timed-out cross-CTA wait (check_fault)
    cudaStream_t copy_stream = nullptr;   // host -> device column chunks of the host-pointer entry points
    static constexpr int kMaxChunks = 128;
    cudaEvent_t chunk_ev[kMaxChunks] = {};  // chunk c of the matrix has landed (created on first use)
    cudaStream_t drain_stream = nullptr;  // device -> host copies of finished rows (host-pointer getrf, RowDrain)
    cudaEvent_t drain_ev[kMaxChunks] = {};
    // panel exchange workspace (see panel.cu)
    void* panel_ws = nullptr;
    size_t panel_ws_bytes = 0;
    uint32_t panel_seq = 0;
    // generic device scratch grown on demand
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    // grow-only work buffers with fixed roles (ensure_work): owned by the context so shutdown releases them
    enum WorkSlot { kWorkCxPackA0 = 0, kWorkCxPackB0, kWorkCxPackA1, kWorkCxPackB1, kWorkQr, kWorkLarf0, kWorkLarf1, kWorkTf32Main, kWorkTf32Aux, kWorkTf32Other, kWorkSlots };
    void* work[kWorkSlots] = {};
    size_t work_bytes[kWorkSlots] = {};
    Options opt;
};

Context& ctx();
// Bumped every time the context is bound to a device (lair_b200_init after a shutdown): per-device caches held in
// function-local statics (cudaFuncSetAttribute done, occupancy, largest cluster size) are redone when it changes.
uint64_t context_epoch();
inline bool stale_for_context(uint64_t& seen) {
    const uint64_t e = context_epoch();
    if (seen == e) return false;
    seen = e;
    return true;
}
// Modules that own device buffers outside Context register a hook; lair_b200_shutdown() runs them (under the call lock)
// so a later lair_b200_init(other device) never dereferences a pointer into the old device's memory.
void register_reset_hook(void (*fn)());
std::mutex& host_call_mutex();  // the lock the host-pointer entry points hold (capi.cu)
// Per-kernel launch configuration cached in a launcher's function-local static: resident blocks per SM (the same on
// every B200) and the devices on which the kernel's attributes have been set (cudaFuncSetAttribute is per device; the
// single-process multi-GPU batched entry launches the same kernel on several).
struct KernCfg {
    int bps = 0;
    unsigned devmask = 0;
    uint64_t epoch = 0;
};
struct ResetHook {
    explicit ResetHook(void (*fn)()) { register_reset_hook(fn); }
};
int ensure_init();                                  // binds to the current device if needed
int ensure_scratch(size_t bytes, void** out);       // device scratch >= bytes (may reallocate)
// work buffer `slot` >= bytes; a buffer that has to grow is released after `s` has drained (its users are ordered on s)
int ensure_work(int slot, size_t bytes, void** out, cudaStream_t s);

// ---- kernels / device-resident routines (row-major, leading dimension in elements) --------
template <class T> int getrf_batched_dev(int64_t batch, int64_t n, T* d_a, int32_t* d_ipiv, int32_t* d_info, cudaStream_t s);
template <class T> int getrf_small_dev(int64_t m, int64_t n, T* d_a, int64_t lda, int32_t* d_ipiv, int32_t* d_info, bool std_layout, cudaStream_t s,
                                       int32_t row_base = 0, bool accumulate = false);
// blocked LU for the complex types: recursive panels, laswp on the real view, ZGEMM as one real (DMMA / FFMA) GEMM on
// packed operands (blocked_cx.cu)
// its leaf panel: <= 8 columns on one cluster, slabs in shared memory (panel_cx.cu); ERR_UNSUPPORTED when it does not fit
template <class T> int panel_cx_dev(int64_t rows, int64_t w, T* d_a, int64_t lda, int32_t* d_ipiv, int32_t row_base, int32_t* d_info, bool std_layout, cudaStream_t s);
template <class T> int getrs_blocked_cx_dev(int64_t n, int64_t nrhs, const T* d_lu, int64_t lda, const int32_t* d_ipiv, T* d_b, int64_t ldb, cudaStream_t s);
template <class T> int getrf_blocked_cx_dev(int64_t m, int64_t n, T* d_a, int64_t lda, int32_t* d_ipiv, int32_t* d_info, bool std_layout, cudaStream_t s);
template <class T> int getrs_small_dev(int64_t n, int64_t nrhs, const T* d_lu, int64_t lda, const int32_t* d_ipiv, T* d_b, int64_t ldb, cudaStream_t s);
// Columns of the device matrix that are still arriving (host -> device copies on another stream)
// while the factorization already runs: chunk c = columns [c*chunk, (c+1)*chunk), ready[c] is
// recorded when it has landed.  The sweep restricts its trailing updates to the chunks that have
// joined and brings a late chunk up to date with one laswp + trsm + gemm when it joins.
struct ColumnFeed {
    int64_t chunk = 0;
    int nchunks = 0;
    const cudaEvent_t* ready = nullptr;
};
// Finished rows of the device matrix leave for the host while the sweep still runs (host-pointer getrf with a
// row-contiguous pinned host array): once every column chunk has joined, the rows above the current block are final
// -- later steps only interchange rows below it --, so at the start of a block step the sweep records an event on its
// main stream and queues a strided device -> host copy of the rows finished since the last one on `stream`.
// `drained` ends as the first row the caller still has to fetch.
struct RowDrain {
    void* host = nullptr;        // element (0, 0) of the host array
    int64_t host_rs = 0;         // its row stride in elements (columns are contiguous)
    cudaStream_t stream = nullptr;
    cudaEvent_t* ev = nullptr;   // nev events (created on first use)
    int nev = 0;
    int64_t min_rows = 0;        // rows per copy at least
    int64_t drained = 0;
    int used = 0;
};
template <class T> int getrf_blocked_dev(int64_t m, int64_t n, T* d_a, int64_t lda, int32_t* d_ipiv, int32_t* d_info, cudaStream_t s,
                                         const ColumnFeed* feed = nullptr, RowDrain* drain = nullptr);
template <class T> int getrs_blocked_dev(int64_t n, int64_t nrhs, const T* d_lu, int64_t lda, const int32_t* d_ipiv, T* d_b, int64_t ldb, cudaStream_t s);
template <class T> int laswp_dev(int64_t ncols, T* d_a, int64_t lda, int64_t k0, int64_t k1, const int32_t* d_ipiv, cudaStream_t s);
// whole k <= 256 unit-lower solve of a wide row block in one launch (trsm_strip.cu); ERR_UNSUPPORTED beyond
template <class T> int trsm_strip_dev(int64_t k, int64_t ncols, const T* d_l, int64_t ldl, T* d_b, int64_t ldb, cudaStream_t s);
template <class T> int trsm_lower_unit_dev(int64_t k, int64_t ncols, const T* d_l, int64_t ldl, T* d_b, int64_t ldb, cudaStream_t s);
template <class T> int trsm_upper_dev(int64_t k, int64_t ncols, const T* d_u, int64_t ldu, T* d_b, int64_t ldb, cudaStream_t s);
template <class T> int gemm_minus_dev(int64_t m, int64_t n, int64_t k, const T* d_a, int64_t lda, const T* d_b, int64_t ldb, T* d_c, int64_t ldc, cudaStream_t s);
template <> int gemm_minus_dev<double>(int64_t m, int64_t n, int64_t k, const double* d_a, int64_t lda, const double* d_b, int64_t ldb, double* d_c, int64_t ldc, cudaStream_t s);  // gemm_f64.cu
// factor the (rows x w) panel at d_a (w <= panel width limit) with partial pivoting;
// d_ipiv[0..w) receive row indices relative to d_a's row 0 plus `row_base`.
template <class T> int panel_dev(int64_t rows, int64_t w, T* d_a, int64_t lda, int32_t* d_ipiv, int32_t row_base, int32_t* d_info, int32_t step_base, cudaStream_t s);
template <class T> int panel_max_width(int64_t rows);
// single-cluster DSMEM variant (panel_cluster.cu); LAIR_B200_ERR_UNSUPPORTED when it does not fit
template <class T> int panel_cluster_dev(int64_t rows, int64_t w, T* d_a, int64_t lda, int32_t* d_ipiv, int32_t row_base, int32_t* d_info, int32_t step_base, cudaStream_t s);
int panel_cluster_max_rows();
int panel_cluster_timing(long long* out8, bool clear);
// fused laswp + unit-lower trsm for k <= 128 (laswp_trsm.cu); LAIR_B200_ERR_UNSUPPORTED beyond
template <class T> int laswp_trsm_dev(int64_t ncols, T* d_a, int64_t lda, int64_t k0, int64_t k, const int32_t* d_ipiv, const T* d_l, int64_t ldl, cudaStream_t s,
                                      int64_t kp = 0);
// batched_lu4.cu: rows retire into the output tile, retired lanes are NaN-poisoned (no liveness bookkeeping)
template <class T> int getrf_batched32v4_dev(int64_t batch, T* d_a, int32_t* d_ipiv, int32_t* d_info, int variant, cudaStream_t s);
template <> int getrf_batched32v4_dev<float>(int64_t batch, float* d_a, int32_t* d_ipiv, int32_t* d_info, int variant, cudaStream_t s);
template <> int getrf_batched32v4_dev<double>(int64_t batch, double* d_a, int32_t* d_ipiv, int32_t* d_info, int variant, cudaStream_t s);
// batched_lu4.cu: straight-line column loop, evidence accumulated, anything but the plain case redone by the exact routine
template <class T> int getrf_batched32v6_dev(int64_t batch, T* d_a, int32_t* d_ipiv, int32_t* d_info, int variant, cudaStream_t s);
template <> int getrf_batched32v6_dev<float>(int64_t batch, float* d_a, int32_t* d_ipiv, int32_t* d_info, int variant, cudaStream_t s);
template <> int getrf_batched32v6_dev<double>(int64_t batch, double* d_a, int32_t* d_ipiv, int32_t* d_info, int variant, cudaStream_t s);
// batched_lu5.cu: the pivot row is broadcast by shuffles, retired rows keep their values in registers
template <class T> int getrf_batched32v8_dev(int64_t batch, T* d_a, int32_t* d_ipiv, int32_t* d_info, int variant, cudaStream_t s);
template <> int getrf_batched32v8_dev<float>(int64_t batch, float* d_a, int32_t* d_ipiv, int32_t* d_info, int variant, cudaStream_t s);
template <> int getrf_batched32v8_dev<double>(int64_t batch, double* d_a, int32_t* d_ipiv, int32_t* d_info, int variant, cudaStream_t s);
// batched_lu6.cu: two matrices per warp, one merged winner store for both, straight-line column step
template <class T> int getrf_batched32v9_dev(int64_t batch, T* d_a, int32_t* d_ipiv, int32_t* d_info, int variant, cudaStream_t s);
template <> int getrf_batched32v9_dev<float>(int64_t batch, float* d_a, int32_t* d_ipiv, int32_t* d_info, int variant, cudaStream_t s);
template <> int getrf_batched32v9_dev<double>(int64_t batch, double* d_a, int32_t* d_ipiv, int32_t* d_info, int variant, cudaStream_t s);
// f32 C -= A B on tcgen05 tensor cores with 3xTF32-split operands (gemm_tf32.cu)
bool sgemm_tf32x3_supported(int64_t m, int64_t n, int64_t k, const float* d_c, int64_t ldc);
int sgemm_tf32x3_minus_dev(int64_t m, int64_t n, int64_t k, const float* d_a, int64_t lda, const float* d_b, int64_t ldb, float* d_c, int64_t ldc,
                           int slot, cudaStream_t s);
// factor one block column stored at local columns [c0, c0+w), diagonal at row r0 (blocked.cu)
template <class T> int getrf_block_dev(int64_t m, T* d_a, int64_t lda, int64_t r0, int64_t c0, int64_t w, int32_t* d_ipiv, int32_t* d_info, cudaStream_t s);
// in-kernel blocked cluster panel (panel_blocked.cu): 8-column register sub-panels, RPT rows per thread
template <class T> int panel_blocked_dev(int64_t rows, int64_t w, T* d_a, int64_t lda, int32_t* d_ipiv, int32_t row_base, int32_t* d_info, int32_t step_base, cudaStream_t s);
template <class T> int panel_blocked_max_width(int64_t rows);
int panel_blocked_timing(long long* out8, bool clear);
// fourth generation (panel_push.cu): the pushed record carries the candidate row's register window, no row pull
template <class T> int panel_push_dev(int64_t rows, int64_t w, T* d_a, int64_t lda, int32_t* d_ipiv, int32_t row_base, int32_t* d_info, int32_t step_base, cudaStream_t s);
template <class T> int panel_push_max_width(int64_t rows);
int panel_push_timing(long long* out8, bool clear);
// dst[r] = final position of the row that starts at r after the interchanges ipiv[0..k1) (laswp_perm.cu)
int laswp_follow_dev(int64_t nrows, int64_t k1, const int32_t* d_ipiv, int32_t* d_dst, cudaStream_t s);
// lu::Factorized::{l, u, p, into_pl} from device-resident factors (lu_extract.cu)
template <class T> int lu_extract_dev(int mode, int64_t m, int64_t n, const T* d_lu, int64_t ld, const int32_t* d_dst, T* d_out, int64_t ldo, cudaStream_t s);
// all interchanges ipiv[k0..k1) on rows [k0, nrows) of a tall narrow matrix, as one permutation (laswp_perm.cu)
template <class T> int laswp_perm_dev(int64_t nrows, int64_t ncols, T* d_a, int64_t lda, int64_t k0, int64_t k1, const int32_t* d_ipiv, cudaStream_t s);
int dtrsm_ll_dev(bool upper, int64_t n, int64_t nrhs, const double* d_lu, int64_t lda, double* d_b, int64_t ldb, cudaStream_t s);
// The kernels that wait on other CTAs through global memory (tall-panel exchange, dataflow solves)
// bound their spins so a lost peer cannot hang the GPU; a spin that runs out raises the context's
// device fault word (Context::d_fault).  check_fault waits for `s`, reads the word and turns a raised
// fault into LAIR_B200_ERR_CUDA with a message: such results are invalid and must not be used.
int check_fault(cudaStream_t s);

// Householder QR (qr.cu): geqrf.rs:9-30 in place on a row-major device matrix, tau on the device;
// qr::Factorized::q (qr.rs:27-59) into a dense m x m device matrix
template <class T> int geqrf_dev(int64_t m, int64_t n, T* d_a, int64_t lda, T* d_tau, cudaStream_t s);
template <class T> int geqrf_unblocked_dev(int64_t m, int64_t n, T* d_a, int64_t lda, T* d_tau, cudaStream_t s);  // one reflector at a time
template <class R> int geqrf_blocked_dev(int64_t m, int64_t n, R* d_a, int64_t lda, R* d_tau, cudaStream_t s);  // f32 / f64 (qr_blocked.cu)
template <class R> int qr_q_blocked_dev(int64_t m, int64_t n, const R* d_qr, int64_t ldqr, const R* d_tau, R* d_q, int64_t ldq, cudaStream_t s);
template <class T> int qr_q_dev(int64_t m, int64_t n, const T* d_qr, int64_t ldqr, const T* d_tau, T* d_q, int64_t ldq, cudaStream_t s);

// ---- dispatch helpers ----------------------------------------------------------------------
template <class T> int getrf_dev(int64_t m, int64_t n, T* d_a, int64_t lda, int32_t* d_ipiv, int32_t* d_info, bool std_layout, cudaStream_t s,
                                 const ColumnFeed* feed = nullptr, RowDrain* drain = nullptr);
template <class T> int getrs_dev(int64_t n, int64_t nrhs, const T* d_lu, int64_t lda, const int32_t* d_ipiv, T* d_b, int64_t ldb, cudaStream_t s);

}  // namespace lair
