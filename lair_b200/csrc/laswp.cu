// lapack::laswp (src/lapack/laswp.rs:11-40) on a device-resident row-major matrix:
// apply the sequential row interchanges (i <-> ipiv[i], i = k0..k1-1) to `ncols` columns.
//
// HBM-bound.  The sequential interchanges are first collapsed into their net effect: every
// thread follows one touched row through the <= 128 transpositions of a pass (pure index
// work in shared memory), which yields a list of (src -> dst) row moves; the moves are then
// executed as a gather into shared memory followed by a scatter.  A CTA owns a strip of 128
// bytes per row (8 lanes of 16 bytes: one full line per row segment, four rows per warp
// instruction), so a pass stages at most 32 KB and a wide range gives hundreds of CTAs
// (round 2: the first version owned 512-byte strips with 128 KB of staging -- one CTA of 8 warps
// per SM, 128 CTAs for 16 384 f32 columns -- and moved 240 GB/s; profiles/r2v_probe_strip.jsonl).
// Algorithmic traffic: 2 * moved_rows * ncols * sizeof(T) bytes (each moved row read once and
// written once).
#include "common.cuh"

namespace lair {
namespace {

constexpr int LASWP_KMAX = 128;  // pivots collapsed per pass (2*KMAX rows staged in smem)
constexpr int LASWP_THREADS = 256;

template <class T, int VEC> struct VecT;
template <> struct VecT<float, 4> { using type = float4; };
template <> struct VecT<double, 2> { using type = double2; };
template <> struct VecT<float, 1> { using type = float; };
template <> struct VecT<double, 1> { using type = double; };

template <class T, int VEC>
__global__ void __launch_bounds__(LASWP_THREADS)
laswp_kernel(T* __restrict__ A, long long lda, int ncols, int k0, int k1, const int32_t* __restrict__ ipiv) {
    using V = typename VecT<T, VEC>::type;
    constexpr int LPR = (VEC > 1) ? 8 : 32;  // lanes per row segment: 8 x 16 bytes, or 32 scalars
    constexpr int RPW = 32 / LPR;            // rows per warp instruction
    __shared__ int s_piv[LASWP_KMAX];
    __shared__ int s_src[2 * LASWP_KMAX];
    __shared__ int s_dst[2 * LASWP_KMAX];
    __shared__ int s_cnt;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    V* buf = reinterpret_cast<V*>(smem_raw);  // [moves][LPR]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = LASWP_THREADS / 32;
    const int sub = lane / LPR, ln = lane % LPR;
    const int col = (blockIdx.x * LPR + ln) * VEC;
    const bool col_ok = col + VEC <= ncols;

    for (int kb = k0; kb < k1; kb += LASWP_KMAX) {
        const int kc = (k1 - kb) < LASWP_KMAX ? (k1 - kb) : LASWP_KMAX;
        if (tid < kc) s_piv[tid] = ipiv[kb + tid];
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        // ---- collapse the kc transpositions (kb+i <-> s_piv[i]) into row moves ----
        if (tid < 2 * kc) {
            int src = -1;
            if (tid < kc) {
                src = kb + tid;  // the kc "top" rows are always touched
            } else {
                const int u = tid - kc;
                const int r = s_piv[u];
                if (r >= kb + kc) {  // a far row; take it once (first occurrence)
                    bool dup = false;
                    for (int v = 0; v < u; ++v) dup |= (s_piv[v] == r);
                    if (!dup) src = r;
                }
            }
            if (src >= 0) {
                int cur = src;
                for (int i = 0; i < kc; ++i) {
                    const int ri = kb + i, p = s_piv[i];
                    if (cur == ri) cur = p;
                    else if (cur == p) cur = ri;
                }
                if (cur != src) {
                    int e = atomicAdd(&s_cnt, 1);
                    s_src[e] = src;
                    s_dst[e] = cur;
                }
            }
        }
        __syncthreads();
        const int nmov = s_cnt;
        if (col_ok) {
#pragma unroll 4
            for (int e = warp * RPW + sub; e < nmov; e += NW * RPW)
                buf[e * LPR + ln] = *reinterpret_cast<const V*>(A + (long long)s_src[e] * lda + col);
        }
        __syncthreads();
        if (col_ok) {
#pragma unroll 4
            for (int e = warp * RPW + sub; e < nmov; e += NW * RPW)
                *reinterpret_cast<V*>(A + (long long)s_dst[e] * lda + col) = buf[e * LPR + ln];
        }
        __syncthreads();
    }
}

template <class T, int VEC>
int launch_laswp(int ncols, T* d_a, int64_t lda, int k0, int k1, const int32_t* d_ipiv, cudaStream_t s) {
    if (ncols <= 0) return LAIR_B200_OK;
    using V = typename VecT<T, VEC>::type;
    constexpr int LPR = (VEC > 1) ? 8 : 32;
    auto kern = laswp_kernel<T, VEC>;
    int kc = (k1 - k0) < LASWP_KMAX ? (k1 - k0) : LASWP_KMAX;
    size_t smem = (size_t)2 * kc * LPR * sizeof(V);  // <= 32 KB on the 16-byte path, 64 KB for unaligned f64
    static bool configured = false;
    static uint64_t seen_epoch = 0;
    if (stale_for_context(seen_epoch)) configured = false;
    if (!configured) {  // (always: the kernel's static tables come on top of the staging, 48 KB is not the threshold to test)
        const size_t maxb = (size_t)2 * LASWP_KMAX * LPR * sizeof(V);
        LAIR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)maxb));
        configured = true;
    }
    unsigned grid = (unsigned)((ncols + LPR * VEC - 1) / (LPR * VEC));
    // upper bound on moved rows: 2 per interchange, each read once and written once
    ProfScope prof(kProfLaswp, s, 4.0 * (double)(k1 - k0) * (double)ncols * sizeof(T));
    kern<<<grid, LASWP_THREADS, smem, s>>>(d_a, (long long)lda, ncols, k0, k1, d_ipiv);
    LAIR_LAUNCH_CHECK();
    return LAIR_B200_OK;
}

}  // namespace

template <class T>
int laswp_dev(int64_t ncols, T* d_a, int64_t lda, int64_t k0, int64_t k1, const int32_t* d_ipiv, cudaStream_t s) {
    LAIR_REQUIRE(ncols >= 0 && k0 >= 0 && k1 >= k0, "laswp: bad range");
    LAIR_REQUIRE(ncols < (1ll << 31) && k1 < (1ll << 31), "laswp: dimension too large");
    if (ncols == 0 || k1 == k0) return LAIR_B200_OK;
    constexpr int VEC = 16 / sizeof(T);
    const bool aligned = (lda % VEC == 0) && (reinterpret_cast<uintptr_t>(d_a) % 16 == 0);
    if (aligned) {
        const int64_t main_cols = ncols / VEC * VEC;
        LAIR_CHECK((launch_laswp<T, VEC>((int)main_cols, d_a, lda, (int)k0, (int)k1, d_ipiv, s)));
        if (main_cols < ncols)
            LAIR_CHECK((launch_laswp<T, 1>((int)(ncols - main_cols), d_a + main_cols, lda, (int)k0, (int)k1, d_ipiv, s)));
        return LAIR_B200_OK;
    }
    return launch_laswp<T, 1>((int)ncols, d_a, lda, (int)k0, (int)k1, d_ipiv, s);
}

template int laswp_dev<float>(int64_t, float*, int64_t, int64_t, int64_t, const int32_t*, cudaStream_t);
template int laswp_dev<double>(int64_t, double*, int64_t, int64_t, int64_t, const int32_t*, cudaStream_t);

}  // namespace lair
