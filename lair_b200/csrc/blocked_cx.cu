// Blocked LU with partial pivoting for the complex scalar types (Complex<f32>, Complex<f64>) -- SURVEY 8f rank 3.
//
// Same shape as the reference's recursive variant (src/lapack/getrf.rs:216-322: factor left / laswp / trsm / gemm /
// factor right / laswp) around an iterative right-looking sweep, like blocked.cu for the real types.  What is
// specific to complex:
//   * The contraction A22 -= L21 * U12 (src/blas/gemm.rs:6-32 on Complex) is ONE real GEMM on packed operands, so it
//     runs on the tuned DMMA (f64) / FFMA (f32) kernels of gemm_f64.cu / gemm.cu with the 4-multiplication flop count:
//         C viewed as real m x 2n  -=  [Re L | Im L] (m x 2k)  *  [ B viewed as real ; B~ ] (2k x 2n),
//         B~[p, 2j] = -Im B[p, j],  B~[p, 2j+1] = Re B[p, j].
//     The two packs move O(mk + kn) elements per O(mnk) update.
//   * Row interchanges do not look at the values: laswp runs on the real view (Complex<f64> = twice the columns of
//     f64, Complex<f32> = one 8-byte word per element) through the vectorised kernel of laswp.cu.
//   * The leaf panel (kLeaf columns, all remaining rows) is factored by one thread-block cluster with the rows in
//     shared memory (panel_cx.cu; the exact single-CTA kernel of small_lu.cu when it does not fit): the
//     reference's loop operation for operation (iamax on |re| + |im|, src/blas/iamax.rs:6-21; Complex reciprocal and
//     products as num-complex 0.4 writes them), so pivot choices inside a leaf see exactly the reference's arithmetic.
//   * The unit-lower solve of U12 works on 32-row triangles in shared memory with the exact (unfused) complex
//     operations; larger triangles recurse through the packed GEMM.
// Parity bar (tests/test_gpu_parity.py::test_complex_blocked_*): pivots identical to the oracle, L\U to rounding,
// scaled backward error <= 10x the oracle's own.
#include <algorithm>

#include "common.cuh"

namespace lair {
namespace {

constexpr int kLeaf = 8;       // leaf panel width
constexpr int kTri = 32;       // triangle solved by one trsm launch
template <class T> struct TrsmCols { static constexpr int value = sizeof(T) == 16 ? 32 : 64; };  // right-hand-side columns per CTA

template <class T> struct RealOf;
template <> struct RealOf<cxf> { using type = float; };
template <> struct RealOf<cxd> { using type = double; };

// ---- operand packs for the real GEMM -----------------------------------------------------------
// Ap (m x 2k, ld 2k): [Re A | Im A]
template <class R>
__global__ void pack_a_kernel(const cx<R>* __restrict__ A, long long lda, int m, int k, R* __restrict__ Ap) {
    const long long total = (long long)m * k;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long i = idx / k;
        const int p = (int)(idx - i * k);
        const cx<R> v = A[i * lda + p];
        Ap[i * (2ll * k) + p] = v.re;
        Ap[i * (2ll * k) + k + p] = v.im;
    }
}
// Bp (2k x 2n, ld 2n): rows 0..k = B as stored (re, im interleaved); rows k..2k = (-im, re)
template <class R>
__global__ void pack_b_kernel(const cx<R>* __restrict__ B, long long ldb, int k, int n, cx<R>* __restrict__ Bp) {
    const long long total = (long long)k * n;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long p = idx / n;
        const int j = (int)(idx - p * n);
        const cx<R> v = B[p * ldb + j];
        Bp[p * (long long)n + j] = v;
        Bp[(p + k) * (long long)n + j] = cx<R>{-v.im, v.re};
    }
}

// grow-only device buffers of the packed operands, owned by the context (one pair per stream role: the caller's stream
// and the lookahead stream run concurrently; every use is ordered on its stream)
int pack_buffers(size_t a_bytes, size_t b_bytes, void** pa, void** pb, cudaStream_t s) {
    const bool aux = s == ctx().aux_stream;
    LAIR_CHECK(ensure_work(aux ? Context::kWorkCxPackA1 : Context::kWorkCxPackA0, a_bytes, pa, s));
    return ensure_work(aux ? Context::kWorkCxPackB1 : Context::kWorkCxPackB0, b_bytes, pb, s);
}

// C (m x n) -= A (m x k) * B (k x n), complex, row-major
template <class T>
int gemm_minus_cx(int64_t m, int64_t n, int64_t k, const T* d_a, int64_t lda, const T* d_b, int64_t ldb, T* d_c, int64_t ldc, cudaStream_t s) {
    using R = typename RealOf<T>::type;
    if (m <= 0 || n <= 0 || k <= 0) return LAIR_B200_OK;
    void *pa, *pb;
    LAIR_CHECK(pack_buffers((size_t)m * 2 * k * sizeof(R), (size_t)2 * k * n * sizeof(T), &pa, &pb, s));
    const int threads = 256;
    auto blocks = [&](int64_t total) { return (unsigned)std::min<int64_t>((total + threads - 1) / threads, (int64_t)ctx().sm_count * 16); };
    pack_a_kernel<R><<<blocks(m * k), threads, 0, s>>>(d_a, (long long)lda, (int)m, (int)k, static_cast<R*>(pa));
    LAIR_LAUNCH_CHECK();
    pack_b_kernel<R><<<blocks(k * n), threads, 0, s>>>(d_b, (long long)ldb, (int)k, (int)n, static_cast<T*>(pb));
    LAIR_LAUNCH_CHECK();
    return gemm_minus_dev<R>(m, 2 * n, 2 * k, static_cast<const R*>(pa), 2 * k, static_cast<const R*>(pb), 2 * n,
                             reinterpret_cast<R*>(d_c), 2 * ldc, s);
}

// ---- laswp on the real view ----------------------------------------------------------------------
int laswp_cx(int64_t ncols, cxd* d_a, int64_t lda, int64_t k0, int64_t k1, const int32_t* ipiv, cudaStream_t s) {
    return laswp_dev<double>(2 * ncols, reinterpret_cast<double*>(d_a), 2 * lda, k0, k1, ipiv, s);
}
int laswp_cx(int64_t ncols, cxf* d_a, int64_t lda, int64_t k0, int64_t k1, const int32_t* ipiv, cudaStream_t s) {
    // one 8-byte word per element: a bit copy, never an arithmetic use of the punned value
    return laswp_dev<double>(ncols, reinterpret_cast<double*>(d_a), lda, k0, k1, ipiv, s);
}

// ---- unit-lower solve, k <= 32: B (k x ncols) <- L^-1 B (src/blas/trsm.rs:6-22) --------------------
template <class T>
__global__ void __launch_bounds__(TrsmCols<T>::value)
trsm_cx32_kernel(const T* __restrict__ L, long long ldl, T* __restrict__ B, long long ldb, int k, int ncols) {
    using O = Ops<T>;
    constexpr int kTrsmCols = TrsmCols<T>::value;
    __shared__ T sl[kTri][kTri + 1];
    __shared__ T sb[kTri][kTrsmCols + 1];
    const int tid = threadIdx.x;
    const int c0 = blockIdx.x * kTrsmCols;
    const int nc = min(kTrsmCols, ncols - c0);
    for (int idx = tid; idx < k * k; idx += kTrsmCols) {
        const int r = idx / k, c = idx - r * k;
        if (c < r) sl[r][c] = L[(long long)r * ldl + c];
    }
    for (int r = 0; r < k; ++r)
        if (tid < nc) sb[r][tid] = B[(long long)r * ldb + c0 + tid];
    __syncthreads();
    if (tid < nc) {
        for (int kk = 0; kk + 1 < k; ++kk) {
            const T bk = sb[kk][tid];
            for (int i = kk + 1; i < k; ++i) sb[i][tid] = O::sub(sb[i][tid], O::mul(sl[i][kk], bk));
        }
    }
    __syncthreads();
    for (int r = 1; r < k; ++r)
        if (tid < nc) B[(long long)r * ldb + c0 + tid] = sb[r][tid];
}

template <class T>
int trsm_lower_unit_cx(int64_t k, int64_t ncols, const T* d_l, int64_t ldl, T* d_b, int64_t ldb, cudaStream_t s) {
    if (k <= 1 || ncols <= 0) return LAIR_B200_OK;
    if (k <= kTri) {
        constexpr int kTrsmCols = TrsmCols<T>::value;
        const unsigned grid = (unsigned)((ncols + kTrsmCols - 1) / kTrsmCols);
        trsm_cx32_kernel<T><<<grid, kTrsmCols, 0, s>>>(d_l, (long long)ldl, d_b, (long long)ldb, (int)k, (int)ncols);
        LAIR_LAUNCH_CHECK();
        return LAIR_B200_OK;
    }
    int64_t k1 = (k / 2) / kTri * kTri;
    if (k1 < kTri) k1 = kTri;
    LAIR_CHECK(trsm_lower_unit_cx<T>(k1, ncols, d_l, ldl, d_b, ldb, s));
    LAIR_CHECK(gemm_minus_cx<T>(k - k1, ncols, k1, d_l + k1 * ldl, ldl, d_b, ldb, d_b + k1 * ldb, ldb, s));  // B2 -= L21 B1
    return trsm_lower_unit_cx<T>(k - k1, ncols, d_l + k1 * ldl + k1, ldl, d_b + k1 * ldb, ldb, s);
}

// ---- upper solve with a true divide by the diagonal, k <= 32: B <- U^-1 B (src/lapack/getrs.rs:30-36) ------------
template <class T>
__global__ void __launch_bounds__(TrsmCols<T>::value)
trsm_upper_cx32_kernel(const T* __restrict__ U, long long ldu, T* __restrict__ B, long long ldb, int k, int ncols) {
    using O = Ops<T>;
    constexpr int kTrsmCols = TrsmCols<T>::value;
    __shared__ T su[kTri][kTri + 1];
    __shared__ T sb[kTri][kTrsmCols + 1];
    const int tid = threadIdx.x;
    const int c0 = blockIdx.x * kTrsmCols;
    const int nc = min(kTrsmCols, ncols - c0);
    for (int idx = tid; idx < k * k; idx += kTrsmCols) {
        const int r = idx / k, c = idx - r * k;
        if (c >= r) su[r][c] = U[(long long)r * ldu + c];
    }
    for (int r = 0; r < k; ++r)
        if (tid < nc) sb[r][tid] = B[(long long)r * ldb + c0 + tid];
    __syncthreads();
    if (tid < nc) {
        for (int i = k - 1; i >= 0; --i) {
            T x = sb[i][tid];
            for (int c = i + 1; c < k; ++c) x = O::sub(x, O::mul(su[i][c], sb[c][tid]));  // k increasing, as the reference
            sb[i][tid] = O::div(x, su[i][i]);
        }
    }
    __syncthreads();
    for (int r = 0; r < k; ++r)
        if (tid < nc) B[(long long)r * ldb + c0 + tid] = sb[r][tid];
}

template <class T>
int trsm_upper_cx(int64_t k, int64_t ncols, const T* d_u, int64_t ldu, T* d_b, int64_t ldb, cudaStream_t s) {
    if (k <= 0 || ncols <= 0) return LAIR_B200_OK;
    if (k <= kTri) {
        constexpr int kTrsmCols = TrsmCols<T>::value;
        const unsigned grid = (unsigned)((ncols + kTrsmCols - 1) / kTrsmCols);
        trsm_upper_cx32_kernel<T><<<grid, kTrsmCols, 0, s>>>(d_u, (long long)ldu, d_b, (long long)ldb, (int)k, (int)ncols);
        LAIR_LAUNCH_CHECK();
        return LAIR_B200_OK;
    }
    int64_t k1 = (k / 2) / kTri * kTri;
    if (k1 < kTri) k1 = kTri;
    // the bottom block first, then eliminate it from the top block: B1 -= U12 X2
    LAIR_CHECK(trsm_upper_cx<T>(k - k1, ncols, d_u + k1 * ldu + k1, ldu, d_b + k1 * ldb, ldb, s));
    LAIR_CHECK(gemm_minus_cx<T>(k1, ncols, k - k1, d_u + k1, ldu, d_b + k1 * ldb, ldb, d_b, ldb, s));
    return trsm_upper_cx<T>(k1, ncols, d_u, ldu, d_b, ldb, s);
}

__global__ void set_info_kernel(int32_t* p, int32_t v) { *p = v; }

template <class T>
struct FactorCx {
    T* A;
    int64_t lda, m, n;
    int32_t* ipiv;
    int32_t* info;
    bool std_layout;
    cudaStream_t s;

    T* at(int64_t r, int64_t c) const { return A + r * lda + c; }
    int swap_cols(int64_t c0, int64_t c1, int64_t k0, int64_t k1, cudaStream_t st) const {
        if (c1 <= c0 || k1 <= k0) return LAIR_B200_OK;
        return laswp_cx(c1 - c0, A + c0, lda, k0, k1, ipiv, st);
    }
    // factor columns [j0, j0 + w) on rows j0..m; pivots land in ipiv[j0 .. j0 + min(w, m - j0))
    int rec(int64_t j0, int64_t w, cudaStream_t st) const {
        const int64_t rows = m - j0;
        if (rows <= 0 || w <= 0) return LAIR_B200_OK;
        if (w <= kLeaf) {
            if (ctx().opt.cx_blocked == 1) {  // one cluster, the panel in shared memory (panel_cx.cu)
                const int rc = panel_cx_dev<T>(rows, w, at(j0, j0), lda, ipiv + j0, (int32_t)j0, info, std_layout, st);
                if (rc != LAIR_B200_ERR_UNSUPPORTED) return rc;
            }
            return getrf_small_dev<T>(rows, w, at(j0, j0), lda, ipiv + j0, info, std_layout, st, (int32_t)j0, true);
        }
        int64_t w1 = (w / 2 + kLeaf - 1) / kLeaf * kLeaf;
        if (w1 >= w) w1 = w - kLeaf;
        LAIR_CHECK(rec(j0, w1, st));
        const int64_t kd = w1 < rows ? w1 : rows;  // pivots produced by the left half
        const int64_t r1 = j0 + kd, c1 = j0 + w1, w2 = w - w1;
        LAIR_CHECK(swap_cols(c1, c1 + w2, j0, r1, st));                                           // laswp (getrf.rs:270-277)
        LAIR_CHECK(trsm_lower_unit_cx<T>(kd, w2, at(j0, j0), lda, at(j0, c1), lda, st));           // trsm  (:278-283)
        if (m > r1 && kd == w1) {
            LAIR_CHECK(gemm_minus_cx<T>(m - r1, w2, w1, at(r1, j0), lda, at(j0, c1), lda, at(r1, c1), lda, st));  // gemm (:289-296)
            LAIR_CHECK(rec(r1, w2, st));                                                             // recurse (:297); r1 == c1
            const int64_t kd2 = (m - r1) < w2 ? (m - r1) : w2;
            LAIR_CHECK(swap_cols(j0, c1, r1, r1 + kd2, st));                                        // laswp left (:308-315)
        }
        return LAIR_B200_OK;
    }
    // trailing update of columns [c0, c1) with the factored block [j0, j0 + jb)
    int update(int64_t j0, int64_t jb, int64_t c0, int64_t c1, cudaStream_t st) const {
        if (c1 <= c0) return LAIR_B200_OK;
        const int64_t r1 = j0 + jb;
        LAIR_CHECK(swap_cols(c0, c1, j0, r1, st));
        LAIR_CHECK(trsm_lower_unit_cx<T>(jb, c1 - c0, at(j0, j0), lda, at(j0, c0), lda, st));
        if (r1 < m) LAIR_CHECK(gemm_minus_cx<T>(m - r1, c1 - c0, jb, at(r1, j0), lda, at(j0, c0), lda, at(r1, c0), lda, st));
        return LAIR_B200_OK;
    }

    // Right-looking sweep with one block of lookahead, as blocked.cu: the panel recursion of block b+1 (stream P, high
    // priority) runs under the bulk of block b's trailing update (stream M).  The two streams touch disjoint column and
    // ipiv ranges; each has its own packed-operand buffers (pack_buffers).
    int run() const {
        const int64_t kmin = m < n ? m : n;
        set_info_kernel<<<1, 1, 0, s>>>(info, -1);
        LAIR_LAUNCH_CHECK();
        const int64_t nb = 128;
        auto width = [&](int64_t j) { return (kmin - j) < nb ? (kmin - j) : nb; };
        const bool look = ctx().opt.lookahead != 0 && kmin > nb;
        cudaStream_t M = s, P = look ? ctx().aux_stream : s;
        cudaEvent_t EP = ctx().ev[0], EN = ctx().ev[1];
        if (look) {
            LAIR_CUDA_CHECK(cudaEventRecord(EN, M));  // P starts after everything already queued on the caller's stream
            LAIR_CUDA_CHECK(cudaStreamWaitEvent(P, EN, 0));
        }
        LAIR_CHECK(rec(0, width(0), P));
        for (int64_t j0 = 0; j0 < kmin; j0 += nb) {
            const int64_t jb = width(j0);
            const int64_t c0 = j0 + jb;
            const int64_t nb2 = c0 < kmin ? width(c0) : 0;
            if (look) {
                LAIR_CUDA_CHECK(cudaEventRecord(EP, P));
                LAIR_CUDA_CHECK(cudaStreamWaitEvent(M, EP, 0));
            }
            if (nb2 > 0) {
                LAIR_CHECK(update(j0, jb, c0, c0 + nb2, M));   // the next block's columns first ...
                if (look) {
                    LAIR_CUDA_CHECK(cudaEventRecord(EN, M));
                    LAIR_CUDA_CHECK(cudaStreamWaitEvent(P, EN, 0));
                }
                LAIR_CHECK(rec(c0, nb2, P));                    // ... so its panel recursion starts under the rest
            }
            LAIR_CHECK(update(j0, jb, c0 + nb2, n, M));
            LAIR_CHECK(swap_cols(0, j0, j0, c0, M));            // interchanges reach back into L (off the critical path)
        }
        if (look) {
            LAIR_CUDA_CHECK(cudaEventRecord(EP, P));
            LAIR_CUDA_CHECK(cudaStreamWaitEvent(M, EP, 0));
        }
        return LAIR_B200_OK;
    }
};

}  // namespace

template <class T>
int getrf_blocked_cx_dev(int64_t m, int64_t n, T* d_a, int64_t lda, int32_t* d_ipiv, int32_t* d_info, bool std_layout, cudaStream_t s) {
    LAIR_REQUIRE(m >= 0 && n >= 0 && lda >= n, "getrf: bad shape m=%lld n=%lld lda=%lld", (long long)m, (long long)n, (long long)lda);
    LAIR_REQUIRE(m < (1ll << 30) && n < (1ll << 30), "getrf: dimension too large");
    if (m == 0 || n == 0) return LAIR_B200_OK;
    FactorCx<T> f{d_a, lda, m, n, d_ipiv, d_info, std_layout, s};
    return f.run();
}
// X = U^-1 L^-1 P B in place in d_b (n x nrhs row-major): getrs.rs:22-36 for every column, blocked.
template <class T>
int getrs_blocked_cx_dev(int64_t n, int64_t nrhs, const T* d_lu, int64_t lda, const int32_t* d_ipiv, T* d_b, int64_t ldb, cudaStream_t s) {
    LAIR_REQUIRE(n >= 0 && nrhs >= 0 && lda >= n && ldb >= nrhs, "getrs: bad shape");
    if (n == 0 || nrhs == 0) return LAIR_B200_OK;
    LAIR_CHECK(laswp_cx(nrhs, d_b, ldb, 0, n, d_ipiv, s));                        // b <- P b (getrs.rs:22-23)
    LAIR_CHECK(trsm_lower_unit_cx<T>(n, nrhs, d_lu, lda, d_b, ldb, s));           // forward  (:24-29)
    return trsm_upper_cx<T>(n, nrhs, d_lu, lda, d_b, ldb, s);                     // backward (:30-36)
}
template int getrs_blocked_cx_dev<cxf>(int64_t, int64_t, const cxf*, int64_t, const int32_t*, cxf*, int64_t, cudaStream_t);
template int getrs_blocked_cx_dev<cxd>(int64_t, int64_t, const cxd*, int64_t, const int32_t*, cxd*, int64_t, cudaStream_t);

template int getrf_blocked_cx_dev<cxf>(int64_t, int64_t, cxf*, int64_t, int32_t*, int32_t*, bool, cudaStream_t);
template int getrf_blocked_cx_dev<cxd>(int64_t, int64_t, cxd*, int64_t, int32_t*, int32_t*, bool, cudaStream_t);

}  // namespace lair
