// lapack::laswp (src/lapack/laswp.rs:11-40) for the right-hand sides of getrs
// (src/lapack/getrs.rs:22): ALL n interchanges applied to a tall, narrow matrix.
//
// laswp.cu walks the pivots 128 at a time inside one CTA per 32-column strip; with n = 8192
// pivots and 64 columns that is 64 dependent passes of a single CTA (0.78 ms measured).  Here
// the sequential interchanges are collapsed into one permutation first -- every row follows
// ITSELF through the whole pivot list (a row that has dropped below the current step can no
// longer move, so most rows stop early) -- and the rows are then moved once through a scratch
// copy.  HBM traffic: 2 reads + 2 writes of the n x ncols matrix (a few MB); the pivot walk is
// index work on a list that stays in shared memory / L1.
#include "common.cuh"

namespace lair {
namespace {

constexpr int LP_THREADS = 256;
constexpr int LP_CHUNK = 2048;  // pivots staged in shared memory per pass

// dst[r] = final position of the row that starts at position r (k0 <= r < nrows)
__global__ void __launch_bounds__(LP_THREADS)
laswp_follow_kernel(int nrows, int k0, int k1, const int32_t* __restrict__ ipiv, int32_t* __restrict__ dst) {
    __shared__ __align__(16) int s_piv[LP_CHUNK];
    const int r = k0 + blockIdx.x * LP_THREADS + threadIdx.x;
    int cur = r;
    for (int base = k0; base < k1; base += LP_CHUNK) {
        const int cnt = (k1 - base) < LP_CHUNK ? (k1 - base) : LP_CHUNK;
        __syncthreads();
        for (int i = threadIdx.x; i < cnt; i += LP_THREADS) s_piv[i] = ipiv[base + i];
        __syncthreads();
        // interchange i swaps positions i and p >= i: a row sitting above i is final (and steps past
        // it are no-ops, so the exit test runs once per four pivots)
        for (int i = 0; i < cnt && base + i <= cur; i += 4) {
            const int4 p4 = *reinterpret_cast<const int4*>(&s_piv[i]);
            const int pv[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (i + e < cnt) {
                    if (cur == base + i + e) cur = pv[e];
                    else if (cur == pv[e]) cur = base + i + e;
                }
            }
        }
    }
    if (r < nrows) dst[r] = cur;
}

template <class T>
__global__ void __launch_bounds__(LP_THREADS)
laswp_scatter_kernel(int nrows, int ncols, int k0, const T* __restrict__ src, long long lds, T* __restrict__ out, long long ldo,
                     const int32_t* __restrict__ dst) {
    // one warp per row, lanes across columns
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = k0 + blockIdx.x * (LP_THREADS / 32) + warp;
    if (r >= nrows) return;
    const int d = dst ? dst[r] : r;
    const T* s = src + (long long)r * lds;
    T* o = out + (long long)d * ldo;
    for (int c = lane; c < ncols; c += 32) o[c] = s[c];
}

struct PermState {
    void* buf = nullptr;
    size_t bytes = 0;
};
PermState g_perm;
void reset_perm_state() {
    if (g_perm.buf) cudaFree(g_perm.buf);
    g_perm = PermState();
}
ResetHook g_perm_hook(reset_perm_state);

}  // namespace

int laswp_follow_dev(int64_t nrows, int64_t k1, const int32_t* d_ipiv, int32_t* d_dst, cudaStream_t s) {
    if (nrows <= 0) return LAIR_B200_OK;
    LAIR_REQUIRE(nrows < (1ll << 31) && k1 >= 0 && k1 <= nrows, "laswp_follow: bad shape");
    const int grid = (int)((nrows + LP_THREADS - 1) / LP_THREADS);
    ProfScope prof(kProfLaswp, s, (double)nrows * 8.0);
    laswp_follow_kernel<<<grid, LP_THREADS, 0, s>>>((int)nrows, 0, (int)k1, d_ipiv, d_dst);
    LAIR_LAUNCH_CHECK();
    return LAIR_B200_OK;
}

// Rows [k0, nrows) of the nrows x ncols matrix d_a take the interchanges ipiv[k0 .. k1).
template <class T>
int laswp_perm_dev(int64_t nrows, int64_t ncols, T* d_a, int64_t lda, int64_t k0, int64_t k1, const int32_t* d_ipiv, cudaStream_t s) {
    LAIR_REQUIRE(nrows >= 0 && ncols >= 0 && k0 >= 0 && k1 >= k0 && k1 <= nrows, "laswp: bad range");
    LAIR_REQUIRE(nrows < (1ll << 31) && ncols < (1ll << 31), "laswp: dimension too large");
    if (ncols == 0 || k1 == k0) return LAIR_B200_OK;
    const int64_t rows = nrows - k0;
    const size_t need = (size_t)nrows * sizeof(int32_t) + 256 + (size_t)nrows * ncols * sizeof(T);
    PermState& st = g_perm;
    if (st.bytes < need) {
        if (st.buf) {
            LAIR_CUDA_CHECK(cudaDeviceSynchronize());
            LAIR_CUDA_CHECK(cudaFree(st.buf));
            st.buf = nullptr;
            st.bytes = 0;
        }
        LAIR_CUDA_CHECK(cudaMalloc(&st.buf, need));
        st.bytes = need;
    }
    int32_t* d_dst = reinterpret_cast<int32_t*>(st.buf);
    T* d_tmp = reinterpret_cast<T*>(reinterpret_cast<char*>(st.buf) + ((size_t)nrows * sizeof(int32_t) + 255) / 256 * 256);
    ProfScope prof(kProfLaswp, s, 4.0 * (double)rows * (double)ncols * sizeof(T));
    const unsigned g1 = (unsigned)((rows + LP_THREADS - 1) / LP_THREADS);
    laswp_follow_kernel<<<g1, LP_THREADS, 0, s>>>((int)nrows, (int)k0, (int)k1, d_ipiv, d_dst);
    LAIR_LAUNCH_CHECK();
    const unsigned g2 = (unsigned)((rows + LP_THREADS / 32 - 1) / (LP_THREADS / 32));
    laswp_scatter_kernel<T><<<g2, LP_THREADS, 0, s>>>((int)nrows, (int)ncols, (int)k0, d_a, (long long)lda, d_tmp, (long long)ncols, d_dst);
    LAIR_LAUNCH_CHECK();
    laswp_scatter_kernel<T><<<g2, LP_THREADS, 0, s>>>((int)nrows, (int)ncols, (int)k0, d_tmp, (long long)ncols, d_a, (long long)lda, nullptr);
    LAIR_LAUNCH_CHECK();
    return LAIR_B200_OK;
}

template int laswp_perm_dev<float>(int64_t, int64_t, float*, int64_t, int64_t, int64_t, const int32_t*, cudaStream_t);
template int laswp_perm_dev<double>(int64_t, int64_t, double*, int64_t, int64_t, int64_t, const int32_t*, cudaStream_t);

}  // namespace lair
