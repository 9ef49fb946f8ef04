// Views of device-resident LU factors: lu::Factorized::{p, l, u, into_pl}
// (src/decomposition/lu.rs:28-72, 107-153) as data-movement kernels on the factors the handle keeps
// in HBM -- the host never sees L\U unless it asks for it.
//   L   (m x k): unit lower triangle of L\U                      lu.rs:42-57
//   U   (k x n): upper triangle of L\U                           lu.rs:60-72
//   P   (m x m): P[perm[i], i] = 1, perm = laswp(identity, piv)  lu.rs:28-39
//   PL  (m x k): rows of L put back in the original row order    lu.rs:107-153
// `dst[r]` = final position of the row that starts at r (laswp_perm.cu's collapsed interchanges);
// perm is its inverse, so P[r, dst[r]] = 1 and PL[r, :] = L[dst[r], :].
// HBM-bound: one read of the triangle + one write of the view.
#include "common.cuh"

namespace lair {
namespace {

template <class T>
__global__ void __launch_bounds__(256)
lu_extract_kernel(int mode, long long rows, long long cols, const T* __restrict__ lu, long long ld, const int32_t* __restrict__ dst,
                  T* __restrict__ out, long long ldo) {
    using O = Ops<T>;
    const long long total = rows * cols;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / cols, c = idx - r * cols;
        T v;
        if (mode == LAIR_LU_VIEW_L) {
            v = c < r ? lu[r * ld + c] : (c == r ? O::one() : O::zero());
        } else if (mode == LAIR_LU_VIEW_U) {
            v = c >= r ? lu[r * ld + c] : O::zero();
        } else if (mode == LAIR_LU_VIEW_P) {
            v = c == (long long)dst[r] ? O::one() : O::zero();
        } else {  // PL
            const long long q = dst[r];
            v = c < q ? lu[q * ld + c] : (c == q ? O::one() : O::zero());
        }
        out[r * ldo + c] = v;
    }
}

}  // namespace

template <class T>
int lu_extract_dev(int mode, int64_t m, int64_t n, const T* d_lu, int64_t ld, const int32_t* d_dst, T* d_out, int64_t ldo, cudaStream_t s) {
    const int64_t k = m < n ? m : n;
    int64_t rows, cols;
    switch (mode) {
        case LAIR_LU_VIEW_L: rows = m; cols = k; break;
        case LAIR_LU_VIEW_U: rows = k; cols = n; break;
        case LAIR_LU_VIEW_P: rows = m; cols = m; break;
        case LAIR_LU_VIEW_PL: rows = m; cols = k; break;
        default: LAIR_REQUIRE(false, "lu view: unknown view %d", mode);
    }
    if (rows == 0 || cols == 0) return LAIR_B200_OK;
    LAIR_REQUIRE(ldo >= cols, "lu view: output leading dimension too small");
    LAIR_REQUIRE((mode != LAIR_LU_VIEW_P && mode != LAIR_LU_VIEW_PL) || d_dst != nullptr, "lu view: permutation missing");
    const long long total = (long long)rows * cols;
    long long want = (total + 255) / 256;
    const long long cap = (long long)ctx().sm_count * 8;
    const int grid = (int)(want < cap ? want : cap);
    ProfScope prof(kProfLaswp, s, (double)total * 2.0 * sizeof(T));
    lu_extract_kernel<T><<<grid, 256, 0, s>>>(mode, rows, cols, d_lu, ld, d_dst, d_out, ldo);
    LAIR_LAUNCH_CHECK();
    return LAIR_B200_OK;
}

template int lu_extract_dev<float>(int, int64_t, int64_t, const float*, int64_t, const int32_t*, float*, int64_t, cudaStream_t);
template int lu_extract_dev<double>(int, int64_t, int64_t, const double*, int64_t, const int32_t*, double*, int64_t, cudaStream_t);
template int lu_extract_dev<cxf>(int, int64_t, int64_t, const cxf*, int64_t, const int32_t*, cxf*, int64_t, cudaStream_t);
template int lu_extract_dev<cxd>(int, int64_t, int64_t, const cxd*, int64_t, const int32_t*, cxd*, int64_t, cudaStream_t);

}  // namespace lair
