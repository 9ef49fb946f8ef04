// Blocked LU with partial pivoting on a device-resident row-major matrix, and the blocked
// multi-RHS solve.  Shape of the algorithm = the reference's recursive variant
// (src/lapack/getrf.rs:216-322: factor left / laswp / trsm / gemm / factor right / laswp),
// which the reference's own test pins to the same pivots and L\U as its live unblocked loops
// (getrf.rs:381-403); here the recursion bottoms out in the on-chip panel kernel (panel.cu)
// instead of a single column, and the outer level is an iterative right-looking sweep of
// width nb so the trailing update is one large DMMA GEMM per block column.
#include "common.cuh"

namespace lair {
namespace {

__global__ void set_i32_kernel(int32_t* p, int32_t v) { *p = v; }

template <class T>
struct Factor {
    T* A;
    int64_t lda, m, n;
    int32_t* ipiv;
    int32_t* info;
    cudaStream_t s;

    T* at(int64_t r, int64_t c) const { return A + r * lda + c; }

    // columns [c0, c1) of the whole matrix get the interchanges ipiv[k0..k1)
    int swap_cols(int64_t c0, int64_t c1, int64_t k0, int64_t k1) const {
        if (c1 <= c0 || k1 <= k0) return LAIR_B200_OK;
        return laswp_dev<T>(c1 - c0, A + c0, lda, k0, k1, ipiv, s);
    }

    // factor columns [j0, j0+w) below (and including) row j0; earlier columns' updates applied
    int rec(int64_t j0, int64_t w) const {
        const int64_t rows = m - j0;
        const int wp = panel_max_width<T>(rows);
        if (wp <= 0) {
            set_error("getrf: %lld rows exceed the on-chip panel capacity", (long long)rows);
            return LAIR_B200_ERR_UNSUPPORTED;
        }
        if (w <= wp) return panel_dev<T>(rows, w, at(j0, j0), lda, ipiv + j0, (int32_t)j0, info, (int32_t)j0, s);
        int64_t w1 = (w / 2 + wp - 1) / wp * wp;  // left half, a multiple of the panel width
        if (w1 >= w) w1 = w - wp > 0 ? (w - 1) / wp * wp : wp;
        LAIR_CHECK(rec(j0, w1));
        const int64_t c0 = j0 + w1, c1 = j0 + w;
        LAIR_CHECK(swap_cols(c0, c1, j0, j0 + w1));                                       // laswp  (getrf.rs:270-277)
        LAIR_CHECK(trsm_lower_unit_dev<T>(w1, c1 - c0, at(j0, j0), lda, at(j0, c0), lda, s));  // trsm   (:278-283)
        if (m > c0)
            LAIR_CHECK(gemm_minus_dev<T>(m - c0, c1 - c0, w1, at(c0, j0), lda, at(j0, c0), lda, at(c0, c0), lda, s));  // gemm (:289-296)
        LAIR_CHECK(rec(c0, c1 - c0));                                                     // recurse (:297)
        return swap_cols(j0, c0, c0, c1);                                                 // laswp left (:308-315)
    }

    int run() const {
        const int64_t kmin = m < n ? m : n;
        set_i32_kernel<<<1, 1, 0, s>>>(info, -1);
        LAIR_LAUNCH_CHECK();
        const int64_t nb = ctx().opt.nb;
        for (int64_t j0 = 0; j0 < kmin; j0 += nb) {
            const int64_t jb = (kmin - j0) < nb ? (kmin - j0) : nb;
            LAIR_CHECK(rec(j0, jb));
            LAIR_CHECK(swap_cols(0, j0, j0, j0 + jb));  // interchanges reach back into L
            const int64_t c0 = j0 + jb;
            if (c0 < n) {
                LAIR_CHECK(swap_cols(c0, n, j0, j0 + jb));
                LAIR_CHECK(trsm_lower_unit_dev<T>(jb, n - c0, at(j0, j0), lda, at(j0, c0), lda, s));
                if (c0 < m)
                    LAIR_CHECK(gemm_minus_dev<T>(m - c0, n - c0, jb, at(c0, j0), lda, at(j0, c0), lda, at(c0, c0), lda, s));
            }
        }
        return LAIR_B200_OK;
    }
};

}  // namespace

template <class T>
int getrf_blocked_dev(int64_t m, int64_t n, T* d_a, int64_t lda, int32_t* d_ipiv, int32_t* d_info, cudaStream_t s) {
    LAIR_REQUIRE(m >= 0 && n >= 0 && lda >= n, "getrf: bad shape m=%lld n=%lld lda=%lld", (long long)m, (long long)n,
                 (long long)lda);
    LAIR_REQUIRE(m < (1ll << 31) && n < (1ll << 31), "getrf: dimension too large");
    if (m == 0 || n == 0) return LAIR_B200_OK;
    Factor<T> f{d_a, lda, m, n, d_ipiv, d_info, s};
    return f.run();
}

// X = U^-1 L^-1 P B, in place in d_b (n x nrhs row-major): getrs.rs:22-36 for every column.
template <class T>
int getrs_blocked_dev(int64_t n, int64_t nrhs, const T* d_lu, int64_t lda, const int32_t* d_ipiv, T* d_b, int64_t ldb,
                      cudaStream_t s) {
    LAIR_REQUIRE(n >= 0 && nrhs >= 0 && lda >= n && ldb >= nrhs, "getrs: bad shape");
    if (n == 0 || nrhs == 0) return LAIR_B200_OK;
    LAIR_CHECK(laswp_dev<T>(nrhs, d_b, ldb, 0, n, d_ipiv, s));
    if constexpr (sizeof(T) == 8) {
        if (ctx().opt.trsm_dataflow) {
            // one persistent dataflow kernel per triangle (trsm_dataflow.cu)
            LAIR_CHECK(dtrsm_dataflow_dev(false, n, nrhs, d_lu, lda, d_b, ldb, s));
            return dtrsm_dataflow_dev(true, n, nrhs, d_lu, lda, d_b, ldb, s);
        }
    }
    LAIR_CHECK(trsm_lower_unit_dev<T>(n, nrhs, d_lu, lda, d_b, ldb, s));
    return trsm_upper_dev<T>(n, nrhs, d_lu, lda, d_b, ldb, s);
}

#define INST(T)                                                                                          \
    template int getrf_blocked_dev<T>(int64_t, int64_t, T*, int64_t, int32_t*, int32_t*, cudaStream_t);   \
    template int getrs_blocked_dev<T>(int64_t, int64_t, const T*, int64_t, const int32_t*, T*, int64_t, cudaStream_t);
INST(float)
INST(double)
#undef INST

}  // namespace lair
