// Triangular solves with many right-hand sides (row-major):
//   trsm_lower_unit:  B <- L^-1 B, L unit lower k x k      (src/blas/trsm.rs:6-22; the U12 step
//                      of the blocked factorization and getrs' forward sweep, getrs.rs:24-29)
//   trsm_upper:       B <- U^-1 B, U upper k x k, true divide by the diagonal (getrs.rs:30-36)
// Blocked recursively on the host: a 32-row base kernel (one thread per right-hand-side
// column, the 32 unknowns in registers, the triangle in shared memory) plus the DMMA GEMM for
// the off-diagonal updates, so almost all flops run on the tensor path.
#include "common.cuh"

namespace lair {
namespace {

constexpr int TB = 32;          // base triangle
constexpr int TRSM_THREADS = 128;

template <class T>
__global__ void __launch_bounds__(TRSM_THREADS)
trsm_lower_unit32_kernel(const T* __restrict__ L, long long ldl, T* __restrict__ B, long long ldb, int k, int ncols) {
    __shared__ T sl[TB][TB + 1];
    const int tid = threadIdx.x;
    for (int idx = tid; idx < TB * TB; idx += TRSM_THREADS) {
        int r = idx / TB, c = idx % TB;
        sl[r][c] = (r < k && c < r) ? L[(long long)r * ldl + c] : T(0);
    }
    __syncthreads();
    const int col = blockIdx.x * TRSM_THREADS + tid;
    if (col >= ncols) return;
    T b[TB];
#pragma unroll
    for (int i = 0; i < TB; ++i) b[i] = (i < k) ? B[(long long)i * ldb + col] : T(0);
#pragma unroll
    for (int kk = 0; kk < TB - 1; ++kk) {
        const T bk = b[kk];
#pragma unroll
        for (int i = kk + 1; i < TB; ++i) b[i] -= sl[i][kk] * bk;
    }
#pragma unroll
    for (int i = 1; i < TB; ++i)
        if (i < k) B[(long long)i * ldb + col] = b[i];
}

template <class T>
__global__ void __launch_bounds__(TRSM_THREADS)
trsm_upper32_kernel(const T* __restrict__ U, long long ldu, T* __restrict__ B, long long ldb, int k, int ncols) {
    __shared__ T su[TB][TB + 1];
    const int tid = threadIdx.x;
    for (int idx = tid; idx < TB * TB; idx += TRSM_THREADS) {
        int r = idx / TB, c = idx % TB;
        T v = T(0);
        if (r < k && c < k && c >= r) v = U[(long long)r * ldu + c];
        if (r >= k && c == r) v = T(1);  // padding rows solve to 0 / 1 = 0
        su[r][c] = v;
    }
    __syncthreads();
    const int col = blockIdx.x * TRSM_THREADS + tid;
    if (col >= ncols) return;
    T b[TB];
#pragma unroll
    for (int i = 0; i < TB; ++i) b[i] = (i < k) ? B[(long long)i * ldb + col] : T(0);
#pragma unroll
    for (int i = TB - 1; i >= 0; --i) {
        b[i] = b[i] / su[i][i];  // true divide, as the reference (getrs.rs:35)
        const T bi = b[i];
#pragma unroll
        for (int r = 0; r < i; ++r) b[r] -= su[r][i] * bi;
    }
#pragma unroll
    for (int i = 0; i < TB; ++i)
        if (i < k) B[(long long)i * ldb + col] = b[i];
}

static int64_t split_point(int64_t k) {
    // largest multiple of 32 that is <= k/2, at least 32
    int64_t h = (k / 2) / TB * TB;
    return h < TB ? TB : h;
}

}  // namespace

template <class T>
int trsm_lower_unit_dev(int64_t k, int64_t ncols, const T* d_l, int64_t ldl, T* d_b, int64_t ldb, cudaStream_t s) {
    LAIR_REQUIRE(k >= 0 && ncols >= 0, "trsm: negative dimension");
    if (k <= 1 || ncols == 0) return LAIR_B200_OK;
    if (k <= TB) {
        unsigned grid = (unsigned)((ncols + TRSM_THREADS - 1) / TRSM_THREADS);
        ProfScope prof(kProfTrsm, s, (double)k * (double)k * (double)ncols);
        trsm_lower_unit32_kernel<T><<<grid, TRSM_THREADS, 0, s>>>(d_l, (long long)ldl, d_b, (long long)ldb, (int)k, (int)ncols);
        LAIR_LAUNCH_CHECK();
        return LAIR_B200_OK;
    }
    const int64_t k1 = split_point(k);
    LAIR_CHECK(trsm_lower_unit_dev<T>(k1, ncols, d_l, ldl, d_b, ldb, s));
    // B2 -= L21 * B1
    LAIR_CHECK(gemm_minus_dev<T>(k - k1, ncols, k1, d_l + k1 * ldl, ldl, d_b, ldb, d_b + k1 * ldb, ldb, s));
    return trsm_lower_unit_dev<T>(k - k1, ncols, d_l + k1 * ldl + k1, ldl, d_b + k1 * ldb, ldb, s);
}

template <class T>
int trsm_upper_dev(int64_t k, int64_t ncols, const T* d_u, int64_t ldu, T* d_b, int64_t ldb, cudaStream_t s) {
    LAIR_REQUIRE(k >= 0 && ncols >= 0, "trsm: negative dimension");
    if (k == 0 || ncols == 0) return LAIR_B200_OK;
    if (k <= TB) {
        unsigned grid = (unsigned)((ncols + TRSM_THREADS - 1) / TRSM_THREADS);
        ProfScope prof(kProfTrsm, s, (double)k * (double)k * (double)ncols);
        trsm_upper32_kernel<T><<<grid, TRSM_THREADS, 0, s>>>(d_u, (long long)ldu, d_b, (long long)ldb, (int)k, (int)ncols);
        LAIR_LAUNCH_CHECK();
        return LAIR_B200_OK;
    }
    const int64_t k1 = split_point(k);
    // solve the bottom block first, then eliminate it from the top block
    LAIR_CHECK(trsm_upper_dev<T>(k - k1, ncols, d_u + k1 * ldu + k1, ldu, d_b + k1 * ldb, ldb, s));
    // B1 -= U12 * X2
    LAIR_CHECK(gemm_minus_dev<T>(k1, ncols, k - k1, d_u + k1, ldu, d_b + k1 * ldb, ldb, d_b, ldb, s));
    return trsm_upper_dev<T>(k1, ncols, d_u, ldu, d_b, ldb, s);
}

#define INST(T)                                                                                         \
    template int trsm_lower_unit_dev<T>(int64_t, int64_t, const T*, int64_t, T*, int64_t, cudaStream_t); \
    template int trsm_upper_dev<T>(int64_t, int64_t, const T*, int64_t, T*, int64_t, cudaStream_t);
INST(float)
INST(double)
#undef INST

}  // namespace lair
