// Single-CTA LU / solve for small matrices, all four scalar types, reference-exact order.
//
// One CTA keeps the whole matrix in shared memory (or works in place in global memory
// when it does not fit -- the slow but always-correct path used for complex matrices
// beyond the shared-memory limit) and runs the reference's right-looking loop
// (src/lapack/getrf.rs:46-120): per step an iamax (src/blas/iamax.rs:6-21), a full-row
// swap, a reciprocal, then `l = a*recip`, `a -= l*u` with separately rounded multiply and
// subtract -- so for standard-layout inputs L\U, ipiv and info are bit-identical to the
// reference.  `std_layout == false` selects the singularity test of the reference's other
// body (pivot == 0, src/lapack/getrf.rs:168,190) instead of max_val == 0 (:72).
//
// getrs_small: src/lapack/getrs.rs:12-38 with the same per-element operation order
// (forward: k increasing; backward: k increasing then a true divide), one CTA per RHS.
#include "common.cuh"

namespace lair {
namespace {

constexpr int kSmallThreads = 512;

template <class R>
__device__ __forceinline__ void arg_better(R& v, int& i, R ov, int oi) {
    // first maximum wins: larger value, or equal value at a lower index
    if (ov > v || (ov == v && oi < i)) {
        v = ov;
        i = oi;
    }
}

template <class T>
__global__ void __launch_bounds__(kSmallThreads, 1)
small_lu_kernel(T* __restrict__ A, long long lda, int m, int n, int32_t* __restrict__ ipiv, int32_t* __restrict__ info,
                int use_smem, int ldw_smem, int std_layout, int row_base, int accumulate) {
    using O = Ops<T>;
    using R = typename O::Real;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ R red_val[kSmallThreads / 32];
    __shared__ int red_idx[kSmallThreads / 32];
    __shared__ int s_piv;
    __shared__ R s_max;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    T* W;
    long long ldw;
    if (use_smem) {
        W = reinterpret_cast<T*>(smem_raw);
        ldw = ldw_smem;
        for (int idx = tid; idx < m * n; idx += kSmallThreads) {
            int r = idx / n, c = idx - r * n;
            W[r * ldw + c] = A[r * lda + c];
        }
    } else {
        W = A;
        ldw = lda;
    }
    __syncthreads();

    const int kmin = m < n ? m : n;
    int sing = -1;
    for (int j = 0; j < kmin; ++j) {
        // ---- iamax down column j, rows j..m ----
        R bv = R(0);
        int bi = j;  // all-zero / all-NaN column -> index 0 (iamax.rs:10-11)
        for (int i = j + tid; i < m; i += kSmallThreads) {
            R v = O::abs1(W[i * ldw + j]);
            if (v > bv) {  // strict: NaN never wins, first max within this thread's rows
                bv = v;
                bi = i;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            R ov = __shfl_xor_sync(0xffffffffu, bv, off);
            int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            arg_better(bv, bi, ov, oi);
        }
        if (lane == 0) {
            red_val[warp] = bv;
            red_idx[warp] = bi;
        }
        __syncthreads();
        if (warp == 0) {
            R v = (lane < kSmallThreads / 32) ? red_val[lane] : R(0);
            int i = (lane < kSmallThreads / 32) ? red_idx[lane] : 0x7fffffff;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                R ov = __shfl_xor_sync(0xffffffffu, v, off);
                int oi = __shfl_xor_sync(0xffffffffu, i, off);
                arg_better(v, i, ov, oi);
            }
            if (lane == 0) {
                // a thread that saw nothing > 0 reports (0, its first row); the block result for an
                // all-zero column must be row j, which is the lowest index any thread reports.
                s_piv = (v > R(0)) ? i : j;
                s_max = v;
                ipiv[j] = row_base + ((v > R(0)) ? i : j);
            }
        }
        __syncthreads();
        const int p = s_piv;
        const R maxv = s_max;
        // ---- swap rows j and p over all columns (getrf.rs:62-71) ----
        if (p != j) {
            for (int c = tid; c < n; c += kSmallThreads) {
                T t = W[j * ldw + c];
                W[j * ldw + c] = W[p * ldw + c];
                W[p * ldw + c] = t;
            }
            __syncthreads();
        }
        const T pivot = W[j * ldw + j];
        const bool singular = std_layout ? (maxv == R(0)) : O::is_zero(pivot);
        if (singular) {
            sing = j;
            continue;  // block-uniform
        }
        const T recip = O::recip(pivot);
        for (int i = j + 1 + tid; i < m; i += kSmallThreads) W[i * ldw + j] = O::mul(W[i * ldw + j], recip);
        __syncthreads();
        // ---- rank-1 update of the trailing block ----
        const int tr = m - j - 1, tc = n - j - 1;
        if (tc > 0) {
            for (int idx = tid; idx < tr * tc; idx += kSmallThreads) {
                int i = idx / tc, c = idx - i * tc;
                T* t = &W[(j + 1 + i) * ldw + (j + 1 + c)];
                *t = O::sub(*t, O::mul(W[(j + 1 + i) * ldw + j], W[j * ldw + (j + 1 + c)]));
            }
        }
        __syncthreads();
    }
    if (use_smem) {
        for (int idx = tid; idx < m * n; idx += kSmallThreads) {
            int r = idx / n, c = idx - r * n;
            A[r * lda + c] = W[r * ldw + c];
        }
    }
    // accumulate (a leaf of the blocked complex factorization, blocked_cx.cu): info keeps the LAST zero-pivot step of
    // the whole matrix (getrf.rs:72-73), so a leaf only writes when it met one
    if (tid == 0 && (!accumulate || sing >= 0)) *info = sing >= 0 ? row_base + sing : sing;
}

// One CTA per right-hand side; x lives in shared memory.
template <class T>
__global__ void __launch_bounds__(256, 1)
small_getrs_kernel(const T* __restrict__ LU, long long lda, int n, const int32_t* __restrict__ ipiv, T* __restrict__ B,
                   long long ldb) {
    using O = Ops<T>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* x = reinterpret_cast<T*>(smem_raw);  // n
    T* prod = x + n;                        // n
    const int tid = threadIdx.x, rhs = blockIdx.x;
    for (int i = tid; i < n; i += blockDim.x) x[i] = B[i * ldb + rhs];
    __syncthreads();
    if (tid == 0) {  // laswp(1, x, ..., 0, p) (getrs.rs:23): sequential interchanges
        for (int i = 0; i < n; ++i) {
            int p = ipiv[i];
            if (p != i) {
                T t = x[i];
                x[i] = x[p];
                x[p] = t;
            }
        }
    }
    __syncthreads();
    // forward: x[i] -= a[i,k]*x[k] for k < i, k increasing (getrs.rs:24-29); a column sweep
    // applies the same operations to each x[i] in the same order.
    for (int k = 0; k + 1 < n; ++k) {
        const T xk = x[k];
        for (int i = k + 1 + tid; i < n; i += blockDim.x) x[i] = O::sub(x[i], O::mul(LU[i * lda + k], xk));
        __syncthreads();
    }
    // backward: for i descending, subtract prod_k for k = i+1.. in increasing k, then divide
    // (getrs.rs:30-36).  Products are formed in parallel; the subtraction chain is sequential.
    for (int i = n - 1; i >= 0; --i) {
        for (int k = i + 1 + tid; k < n; k += blockDim.x) prod[k] = O::mul(LU[i * lda + k], x[k]);
        __syncthreads();
        if (tid == 0) {
            T xi = x[i];
            for (int k = i + 1; k < n; ++k) xi = O::sub(xi, prod[k]);
            x[i] = O::div(xi, LU[i * lda + i]);
        }
        __syncthreads();
    }
    for (int i = tid; i < n; i += blockDim.x) B[i * ldb + rhs] = x[i];
}

}  // namespace

template <class T>
int getrf_small_dev(int64_t m, int64_t n, T* d_a, int64_t lda, int32_t* d_ipiv, int32_t* d_info, bool std_layout,
                    cudaStream_t s, int32_t row_base, bool accumulate) {
    LAIR_REQUIRE(m >= 0 && n >= 0 && lda >= n, "getrf_small: bad shape m=%lld n=%lld lda=%lld", (long long)m,
                 (long long)n, (long long)lda);
    LAIR_REQUIRE(m < (1 << 20) && n < (1 << 20) && m * n < (1ll << 31), "getrf_small: matrix too large");
    if (m == 0 || n == 0) {
        int32_t none = -1;
        LAIR_CUDA_CHECK(cudaMemcpyAsync(d_info, &none, sizeof(none), cudaMemcpyHostToDevice, s));
        LAIR_CUDA_CHECK(cudaStreamSynchronize(s));
        return LAIR_B200_OK;
    }
    auto kern = small_lu_kernel<T>;
    // odd leading dimension (in 8-byte words) spreads a column over the banks
    int ldw = (int)n;
    if ((ldw * (int)sizeof(T) / 4) % 2 == 0) ldw += 1;
    size_t need = (size_t)m * ldw * sizeof(T);
    size_t limit = ctx().smem_optin > 4096 ? ctx().smem_optin - 4096 : 0;
    int use_smem = need <= limit;
    size_t smem = use_smem ? need : 0;
    static size_t configured = 0;
    static uint64_t seen_epoch = 0;
    if (stale_for_context(seen_epoch)) configured = 0;
    if (smem > configured) {
        LAIR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
        configured = limit;
    }
    ProfScope prof(kProfSmall, s, 2.0 / 3.0 * (double)m * (double)n * (double)(m < n ? m : n));
    kern<<<1, kSmallThreads, smem, s>>>(d_a, (long long)lda, (int)m, (int)n, d_ipiv, d_info, use_smem, ldw,
                                        std_layout ? 1 : 0, (int)row_base, accumulate ? 1 : 0);
    LAIR_LAUNCH_CHECK();
    return LAIR_B200_OK;
}

template <class T>
int getrs_small_dev(int64_t n, int64_t nrhs, const T* d_lu, int64_t lda, const int32_t* d_ipiv, T* d_b, int64_t ldb,
                    cudaStream_t s) {
    LAIR_REQUIRE(n >= 0 && nrhs >= 0 && lda >= n && ldb >= nrhs, "getrs_small: bad shape");
    if (n == 0 || nrhs == 0) return LAIR_B200_OK;
    auto kern = small_getrs_kernel<T>;
    size_t smem = 2 * (size_t)n * sizeof(T);
    const size_t limit = ctx().smem_optin > 2048 ? ctx().smem_optin - 2048 : 0;
    LAIR_REQUIRE(smem <= limit, "getrs_small: n=%lld too large", (long long)n);
    static bool configured = false;
    static uint64_t seen_epoch2 = 0;
    if (stale_for_context(seen_epoch2)) configured = false;
    if (!configured) {
        LAIR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
        configured = true;
    }
    kern<<<(unsigned)nrhs, 256, smem, s>>>(d_lu, (long long)lda, (int)n, d_ipiv, d_b, (long long)ldb);
    LAIR_LAUNCH_CHECK();
    return LAIR_B200_OK;
}

#define INST(T)                                                                                              \
    template int getrf_small_dev<T>(int64_t, int64_t, T*, int64_t, int32_t*, int32_t*, bool, cudaStream_t, int32_t, bool);  \
    template int getrs_small_dev<T>(int64_t, int64_t, const T*, int64_t, const int32_t*, T*, int64_t, cudaStream_t);
INST(float)
INST(double)
INST(cxf)
INST(cxd)
#undef INST

}  // namespace lair
