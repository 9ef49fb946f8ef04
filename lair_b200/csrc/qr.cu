// Householder QR on a device-resident row-major matrix -- SURVEY 8f rank 4: lapack::geqrf
// (src/lapack/geqrf.rs:9-30) and qr::Factorized::q (src/decomposition/qr.rs:27-59), all four scalar types.
//
// The reference's loop, one reflector at a time:
//   larfg  (src/lapack/larfg.rs:9-42)   one CTA: |x|^2 by a block reduction, then beta / tau / the scale factor exactly as
//                                       the reference forms them (including its safe-minimum rescaling loop, restated
//                                       literally), x scaled in place, beta stored on the diagonal, tau in d_tau;
//   larf::left (src/lapack/larf.rs:10-54)  C := (I - tau v v^H) C on the trailing block: one CTA per 64 columns, rows
//                                       walked by 4 row groups with coalesced row segments; w = C^H v (gemv::conjtrans,
//                                       src/blas/gemv.rs:58-88) then C += (-tau v) w^H (gerc, src/blas/gerc.rs:8-34) with
//                                       v[0] = 1 implicit (the reference writes the 1 and restores beta, geqrf.rs:22-26).
// Both are HBM / L2-bound sweeps (3 passes over the trailing block per reflector); arithmetic is unfused (Ops<T>).
// Differences from the reference that stay within rounding: the sums of nrm2 and of the column dot products are
// formed in parallel order, and the reference's trimming of trailing zero rows / columns (ilalc) is not needed
// (it only skips multiplications by zero).  Parity bar (tests/test_gpu_qr.py): the reference's golden vectors,
// |QR - QR_oracle| and |tau - tau_oracle| to rounding, ||A - QR|| and ||Q^H Q - I|| <= 10x the oracle's own.
#include <algorithm>

#include "common.cuh"

namespace lair {
namespace {

constexpr int QR_THREADS = 256;
constexpr int QR_COLS = 64;
constexpr int QR_ROWG = QR_THREADS / QR_COLS;

// ---- scalar helpers missing from Ops<T>: conjugate, real <-> scalar ---------------------------------
template <class R> __device__ __forceinline__ cx<R> conj_s(cx<R> a) { return {a.re, -a.im}; }
__device__ __forceinline__ float conj_s(float a) { return a; }
__device__ __forceinline__ double conj_s(double a) { return a; }
template <class R> __device__ __forceinline__ R re_s(cx<R> a) { return a.re; }
template <class R> __device__ __forceinline__ R im_s(cx<R> a) { return a.im; }
__device__ __forceinline__ float re_s(float a) { return a; }
__device__ __forceinline__ float im_s(float) { return 0.f; }
__device__ __forceinline__ double re_s(double a) { return a; }
__device__ __forceinline__ double im_s(double) { return 0.0; }
template <class R> __device__ __forceinline__ cx<R> times_real(cx<R> a, R r) { return {Ops<R>::mul(a.re, r), Ops<R>::mul(a.im, r)}; }
__device__ __forceinline__ float times_real(float a, float r) { return __fmul_rn(a, r); }
__device__ __forceinline__ double times_real(double a, double r) { return __dmul_rn(a, r); }
template <class R> __device__ __forceinline__ cx<R> over_real(cx<R> a, R r) { return {Ops<R>::div(a.re, r), Ops<R>::div(a.im, r)}; }
__device__ __forceinline__ float over_real(float a, float r) { return __fdiv_rn(a, r); }
__device__ __forceinline__ double over_real(double a, double r) { return __ddiv_rn(a, r); }
template <class T> struct FromReal;
template <> struct FromReal<float> { __device__ static float make(float r) { return r; } };
template <> struct FromReal<double> { __device__ static double make(double r) { return r; } };
template <class R> struct FromReal<cx<R>> { __device__ static cx<R> make(R r) { return {r, R(0)}; } };
template <class R> struct RealLimits;
template <> struct RealLimits<float> {
    __device__ static float eps() { return 5.9604644775390625e-08f; }      // Real::eps = epsilon / 2 (src/scalar.rs:393-395)
    __device__ static float sfmin() { return 1.17549435082228750797e-38f; }  // min_positive_value (:400-402)
};
template <> struct RealLimits<double> {
    __device__ static double eps() { return 1.1102230246251565404e-16; }
    __device__ static double sfmin() { return 2.2250738585072013831e-308; }
};
template <class T> __device__ __forceinline__ T neg_s(T a) { return Ops<T>::sub(Ops<T>::zero(), a); }

// block-wide sum (parallel order); result valid in every thread
template <class R>
__device__ R block_sum(R v, R* red /* [blockDim.x / 32] */) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();  // red may still be read from a previous call
    if (lane == 0) red[warp] = v;
    __syncthreads();
    R t = R(0);
    for (int w = 0; w < nw; ++w) t += red[w];
    return t;
}

// larfg on column `col`: alpha = A[0, 0] of the view, x = A[1.., 0] (stride lda).  Writes beta to A[0, 0], the scaled x
// in place and tau to *tau_out.
template <class T>
__global__ void __launch_bounds__(512)
larfg_kernel(T* __restrict__ A, long long lda, int rows, T* __restrict__ tau_out) {
    using O = Ops<T>;
    using R = typename O::Real;
    __shared__ R red[16];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int n = rows - 1;
    T* x = A + lda;
    auto sumsq = [&]() {
        R s = R(0);
        for (int i = tid; i < n; i += nt) {
            const T v = x[(long long)i * lda];
            s += re_s(v) * re_s(v) + im_s(v) * im_s(v);
        }
        return block_sum<R>(s, red);
    };
    T alpha = A[0];
    R x_norm = sqrt(sumsq());
    if (x_norm == R(0) && im_s(alpha) == R(0)) {  // larfg.rs:15-17: H = I
        if (tid == 0) *tau_out = O::zero();
        return;
    }
    const R ar = re_s(alpha), ai = im_s(alpha);
    R beta = -copysign(sqrt(ar * ar + ai * ai + x_norm * x_norm), ar);  // lapy3 (src/lapack.rs:63-65)
    const R safe_min = RealLimits<R>::sfmin() / RealLimits<R>::eps();
    int knt = 0;
    if (fabs(beta) < safe_min) {  // block-uniform: every thread holds the same beta
        const R rsm = R(1) / safe_min;
        for (;;) {
            ++knt;
            for (int i = tid; i < n; i += nt) x[(long long)i * lda] = times_real(x[(long long)i * lda], rsm);
            beta *= rsm;
            alpha = times_real(alpha, rsm);
            if (fabs(beta) >= safe_min || knt >= 20) break;
        }
        __syncthreads();
        x_norm = sqrt(sumsq());
        const R asq = re_s(alpha) * re_s(alpha) + im_s(alpha) * im_s(alpha);
        beta = -copysign(asq + x_norm * x_norm, re_s(alpha));  // literal (larfg.rs:33)
    }
    const T beta_t = FromReal<T>::make(beta);
    const T tau = over_real(O::sub(beta_t, alpha), beta);
    const T scale = O::div(O::one(), O::sub(alpha, beta_t));
    for (int i = tid; i < n; i += nt) x[(long long)i * lda] = O::mul(x[(long long)i * lda], scale);
    for (int k = 0; k < knt; ++k) beta *= safe_min;  // beta *= safe_min^knt
    if (tid == 0) {
        A[0] = FromReal<T>::make(beta);
        *tau_out = tau;
    }
}

// C := (I - t v v^H) C with v = [1; V[1.., 0]] (V column stride ldv), t = *tau or conj(*tau); C is rows x ncols.
template <class T>
__global__ void __launch_bounds__(QR_THREADS)
larf_left_kernel(const T* __restrict__ V, long long ldv, const T* __restrict__ tau, int conj_tau, T* __restrict__ C, long long ldc,
                 int rows, int ncols) {
    using O = Ops<T>;
    __shared__ T part[QR_ROWG][QR_COLS];
    T t = *tau;
    if (O::is_zero(t)) return;  // larf.rs:16-18
    if (conj_tau) t = conj_s(t);
    const int tx = threadIdx.x % QR_COLS, ty = threadIdx.x / QR_COLS;
    const int col = blockIdx.x * QR_COLS + tx;
    const bool live = col < ncols;
    // w[col] = sum_r conj(C[r, col]) v[r]
    T acc = O::zero();
    if (live) {
        for (int r = ty; r < rows; r += QR_ROWG) {
            const T v = r == 0 ? O::one() : V[(long long)r * ldv];
            acc = O::add(acc, O::mul(conj_s(C[(long long)r * ldc + col]), v));
        }
    }
    part[ty][tx] = acc;
    __syncthreads();
    T w = part[0][tx];
#pragma unroll
    for (int g = 1; g < QR_ROWG; ++g) w = O::add(w, part[g][tx]);
    const T wc = conj_s(w);
    const T nt = neg_s(t);
    if (live) {
        for (int r = ty; r < rows; r += QR_ROWG) {
            const T v = r == 0 ? O::one() : V[(long long)r * ldv];
            const T factor = O::mul(nt, v);
            T* c = &C[(long long)r * ldc + col];
            *c = O::add(*c, O::mul(factor, wc));
        }
    }
}

// qr::Factorized::q, start (qr.rs:28-39): Q[:, j < k] = QR[:, j], the other columns = e_j
template <class T>
__global__ void q_init_kernel(const T* __restrict__ QR, long long ldqr, int m, int k, T* __restrict__ Q, long long ldq) {
    const long long total = (long long)m * m;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(idx / m), c = (int)(idx - (long long)r * m);
        Q[(long long)r * ldq + c] = c < k ? QR[(long long)r * ldqr + c] : (r == c ? Ops<T>::one() : Ops<T>::zero());
    }
}
// qr.rs:47-57 for reflector i: Q[i+1.., i] *= -tau; Q[i, i] = 1 - tau; Q[..i, i] = 0
template <class T>
__global__ void q_column_kernel(T* __restrict__ Q, long long ldq, int m, int i, const T* __restrict__ tau) {
    using O = Ops<T>;
    const T t = tau[i];
    const T nt = neg_s(t);
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += gridDim.x * blockDim.x) {
        T* q = &Q[(long long)r * ldq + i];
        if (r < i) *q = O::zero();
        else if (r == i) *q = O::sub(O::one(), t);
        else *q = O::mul(*q, nt);
    }
}

// The same reflector application split over row chunks for tall blocks with few columns (a 65 536 x 256 matrix gives
// the column-parallel kernel 4 CTAs): pass 1 writes the partial dot products of every (column tile, row chunk), pass 2
// sums them in chunk order and updates its chunk.
constexpr int QR_RCHUNK = 1024;
template <class T>
__global__ void __launch_bounds__(QR_THREADS)
larf_dot_kernel(const T* __restrict__ V, long long ldv, const T* __restrict__ tau, const T* __restrict__ C, long long ldc, int rows, int ncols,
                T* __restrict__ part /* [chunks][ncols] */) {
    using O = Ops<T>;
    __shared__ T red[QR_ROWG][QR_COLS];
    if (O::is_zero(*tau)) return;
    const int tx = threadIdx.x % QR_COLS, ty = threadIdx.x / QR_COLS;
    const int col = blockIdx.x * QR_COLS + tx;
    const int r0 = blockIdx.y * QR_RCHUNK, r1 = min(rows, r0 + QR_RCHUNK);
    T acc = O::zero();
    if (col < ncols) {
        for (int r = r0 + ty; r < r1; r += QR_ROWG) {
            const T v = r == 0 ? O::one() : V[(long long)r * ldv];
            acc = O::add(acc, O::mul(conj_s(C[(long long)r * ldc + col]), v));
        }
    }
    red[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && col < ncols) {
        T w = red[0][tx];
#pragma unroll
        for (int g = 1; g < QR_ROWG; ++g) w = O::add(w, red[g][tx]);
        part[(long long)blockIdx.y * ncols + col] = w;
    }
}
template <class T>
__global__ void __launch_bounds__(QR_THREADS)
larf_apply_kernel(const T* __restrict__ V, long long ldv, const T* __restrict__ tau, int conj_tau, T* __restrict__ C, long long ldc, int rows,
                  int ncols, const T* __restrict__ part, int nchunks) {
    using O = Ops<T>;
    T t = *tau;
    if (O::is_zero(t)) return;
    if (conj_tau) t = conj_s(t);
    const int tx = threadIdx.x % QR_COLS, ty = threadIdx.x / QR_COLS;
    const int col = blockIdx.x * QR_COLS + tx;
    if (col >= ncols) return;
    T w = O::zero();
    for (int c = 0; c < nchunks; ++c) w = O::add(w, part[(long long)c * ncols + col]);
    const T wc = conj_s(w);
    const T nt = neg_s(t);
    const int r0 = blockIdx.y * QR_RCHUNK, r1 = min(rows, r0 + QR_RCHUNK);
    for (int r = r0 + ty; r < r1; r += QR_ROWG) {
        const T v = r == 0 ? O::one() : V[(long long)r * ldv];
        T* c = &C[(long long)r * ldc + col];
        *c = O::add(*c, O::mul(O::mul(nt, v), wc));
    }
}

// partial-product scratch of the row-split path: one context-owned buffer per stream role (caller's stream / lookahead stream)
template <class T>
int larf_scratch(size_t elems, T** out, cudaStream_t s) {
    void* p = nullptr;
    LAIR_CHECK(ensure_work(s == ctx().aux_stream ? Context::kWorkLarf1 : Context::kWorkLarf0, elems * sizeof(T), &p, s));
    *out = static_cast<T*>(p);
    return LAIR_B200_OK;
}

template <class T>
int larf_left_dev(const T* d_v, int64_t ldv, const T* d_tau, bool conj_tau, T* d_c, int64_t ldc, int64_t rows, int64_t ncols, cudaStream_t s) {
    if (rows <= 0 || ncols <= 0) return LAIR_B200_OK;
    const unsigned grid = (unsigned)((ncols + QR_COLS - 1) / QR_COLS);
    if (rows >= 4 * QR_RCHUNK && (int)grid * 2 <= ctx().sm_count) {  // tall and narrow: split the rows as well
        const int nchunks = (int)((rows + QR_RCHUNK - 1) / QR_RCHUNK);
        T* part = nullptr;
        LAIR_CHECK(larf_scratch<T>((size_t)nchunks * (size_t)ncols, &part, s));
        larf_dot_kernel<T><<<dim3(grid, (unsigned)nchunks), QR_THREADS, 0, s>>>(d_v, (long long)ldv, d_tau, d_c, (long long)ldc, (int)rows, (int)ncols, part);
        LAIR_LAUNCH_CHECK();
        larf_apply_kernel<T><<<dim3(grid, (unsigned)nchunks), QR_THREADS, 0, s>>>(d_v, (long long)ldv, d_tau, conj_tau ? 1 : 0, d_c, (long long)ldc, (int)rows,
                                                                              (int)ncols, part, nchunks);
        LAIR_LAUNCH_CHECK();
        return LAIR_B200_OK;
    }
    larf_left_kernel<T><<<grid, QR_THREADS, 0, s>>>(d_v, (long long)ldv, d_tau, conj_tau ? 1 : 0, d_c, (long long)ldc, (int)rows, (int)ncols);
    LAIR_LAUNCH_CHECK();
    return LAIR_B200_OK;
}

}  // namespace

// geqrf.rs:9-30 on a row-major device matrix; d_tau has min(m, n) entries.
template <class T>
int geqrf_dev(int64_t m, int64_t n, T* d_a, int64_t lda, T* d_tau, cudaStream_t s) {
    LAIR_REQUIRE(m >= 0 && n >= 0 && lda >= n, "geqrf: bad shape m=%lld n=%lld lda=%lld", (long long)m, (long long)n, (long long)lda);
    LAIR_REQUIRE(m < (1ll << 31) && n < (1ll << 31), "geqrf: dimension too large");
    const int64_t k = m < n ? m : n;
    if constexpr (!Ops<T>::is_complex) {
        // f32 / f64 beyond a few panels: compact-WY blocks on the cluster panel + GEMM machinery (qr_blocked.cu)
        if (ctx().opt.qr_blocked != 0 && k >= 64) {
            const int rc = geqrf_blocked_dev<T>(m, n, d_a, lda, d_tau, s);
            if (rc != LAIR_B200_ERR_UNSUPPORTED) return rc;
        }
    }
    return geqrf_unblocked_dev<T>(m, n, d_a, lda, d_tau, s);
}

// the reference's loop, one reflector at a time (also the panel of the blocked sweep when it does not fit one cluster)
template <class T>
int geqrf_unblocked_dev(int64_t m, int64_t n, T* d_a, int64_t lda, T* d_tau, cudaStream_t s) {
    const int64_t k = m < n ? m : n;
    ProfScope prof(kProfSmall, s, 2.0 * (double)m * (double)n * (double)k);
    for (int64_t i = 0; i < k; ++i) {
        T* aii = d_a + i * lda + i;
        larfg_kernel<T><<<1, 512, 0, s>>>(aii, (long long)lda, (int)(m - i), d_tau + i);
        LAIR_LAUNCH_CHECK();
        if (i + 1 < n) LAIR_CHECK(larf_left_dev<T>(aii, lda, d_tau + i, true, aii + 1, lda, m - i, n - i - 1, s));  // geqrf.rs:24
    }
    return LAIR_B200_OK;
}

// qr::Factorized::q (qr.rs:27-59): d_q is m x m (ldq >= m), built from the factored matrix and tau.
template <class T>
int qr_q_dev(int64_t m, int64_t n, const T* d_qr, int64_t ldqr, const T* d_tau, T* d_q, int64_t ldq, cudaStream_t s) {
    LAIR_REQUIRE(m >= 0 && n >= 0 && ldqr >= n && ldq >= m, "qr_q: bad shape");
    if (m == 0) return LAIR_B200_OK;
    const int64_t k = m < n ? m : n;
    if constexpr (!Ops<T>::is_complex) {
        if (ctx().opt.qr_blocked != 0 && k >= 64) return qr_q_blocked_dev<T>(m, n, d_qr, ldqr, d_tau, d_q, ldq, s);
    }
    const unsigned blocks = (unsigned)std::min<int64_t>((m * m + 255) / 256, (int64_t)ctx().sm_count * 8);
    q_init_kernel<T><<<blocks, 256, 0, s>>>(d_qr, (long long)ldqr, (int)m, (int)k, d_q, (long long)ldq);
    LAIR_LAUNCH_CHECK();
    for (int64_t i = k - 1; i >= 0; --i) {
        T* qii = d_q + i * ldq + i;
        if (i + 1 < m) LAIR_CHECK(larf_left_dev<T>(qii, ldq, d_tau + i, false, qii + 1, ldq, m - i, m - i - 1, s));  // qr.rs:43-46
        q_column_kernel<T><<<(unsigned)((m + 255) / 256), 256, 0, s>>>(d_q, (long long)ldq, (int)m, (int)i, d_tau);
        LAIR_LAUNCH_CHECK();
    }
    return LAIR_B200_OK;
}

#define INST(T)                                                                        \
    template int geqrf_dev<T>(int64_t, int64_t, T*, int64_t, T*, cudaStream_t);         \
    template int geqrf_unblocked_dev<T>(int64_t, int64_t, T*, int64_t, T*, cudaStream_t); \
    template int qr_q_dev<T>(int64_t, int64_t, const T*, int64_t, const T*, T*, int64_t, cudaStream_t);
INST(float)
INST(double)
INST(cxf)
INST(cxd)
#undef INST

}  // namespace lair
