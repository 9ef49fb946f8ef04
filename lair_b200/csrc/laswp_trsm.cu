// Fused row interchange + unit-lower triangular solve for one block step of the factorization:
//     A[:, cols]  <-  laswp(ipiv[k0 .. k0+k))           (src/lapack/laswp.rs:11-40)
//     A[k0..k0+k, cols]  <-  L11^-1 * A[k0..k0+k, cols]  (src/blas/trsm.rs:6-22)
// i.e. the laswp + trsm pair of the blocked LU (reference shape: src/lapack/getrf.rs:270-283) in
// ONE launch for k <= 128.  On the lookahead's critical path the two used to be 2-4 dependent
// latency-bound launches (~12 us each); fused, the moved rows are gathered once, the top k rows are
// solved in shared memory, and everything is scattered once.
//
// One CTA per strip of COLS columns.  The k sequential transpositions are collapsed into row moves
// exactly as in laswp.cu; rows that land in the top block [k0, k0+k) are loaded into the `top`
// tile (by destination), rows that leave it into the `far` buffer.  HBM traffic: each touched row
// segment read once and written once.
#include "common.cuh"

namespace lair {
namespace {

constexpr int LT_THREADS = 256;  // >= 2 * KMAX (one thread per candidate row of the collapse)

// LT_KMAX = 64 (33 KB / 67 KB of shared memory, f32 / f64) or 128 (100 KB / 200 KB)
template <class T, int LT_KMAX>
struct LtCfg {
    static constexpr int COLS = 32;             // columns per CTA
    static constexpr int LDT = COLS + 1;        // top / far tile pitch
    static constexpr int LDL = LT_KMAX + 16 / (int)sizeof(T);  // L tile pitch: 16-byte aligned rows for the 128-bit multiplier loads
    static constexpr int PCH = 64;              // prefix rows applied per pass
    static constexpr size_t smem_bytes = (size_t)(2 * LT_KMAX * LDT + LT_KMAX * LDL + PCH * LDT) * sizeof(T);
};

template <class T, int LT_KMAX>
__global__ void __launch_bounds__(LT_THREADS)
laswp_trsm_kernel(T* __restrict__ A, long long lda, int ncols, int k0, int k, const int32_t* __restrict__ ipiv,
                  const T* __restrict__ L, long long ldl, int kp) {
    using C = LtCfg<T, LT_KMAX>;
    constexpr int COLS = C::COLS, LDT = C::LDT, LDL = C::LDL, PCH = C::PCH;
    static_assert(LT_KMAX >= PCH, "the L tile doubles as the prefix chunk buffer");
    __shared__ int s_piv[LT_KMAX];
    __shared__ int s_topsrc[LT_KMAX];        // source row of each top destination row
    __shared__ int s_farsrc[LT_KMAX];        // moves that leave the top block: source (a top row) ...
    __shared__ int s_fardst[LT_KMAX];        // ... and destination (a far row)
    __shared__ int s_nfar;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* top = reinterpret_cast<T*>(smem_raw);     // [k][LDT]
    T* far = top + LT_KMAX * LDT;                // [nfar][LDT]
    T* Ls = far + LT_KMAX * LDT;                 // [k][LDL]
    T* Ub = Ls + LT_KMAX * LDL;                  // [PCH][LDT] rows above the block (prefix update)

    const int tid = threadIdx.x;
    const int col0 = blockIdx.x * COLS;
    const int cw = (ncols - col0) < COLS ? (ncols - col0) : COLS;  // columns of this strip

    if (tid < k) {
        s_piv[tid] = ipiv[k0 + tid];
        s_topsrc[tid] = k0 + tid;
    }
    if (tid == 0) s_nfar = 0;
    // the triangle (strictly lower part of L11; unit diagonal implied); with a prefix the tile is
    // first used for the prefix chunks of L and the triangle is loaded afterwards
    auto load_triangle = [&]() {
        for (int idx = tid; idx < k * k; idx += LT_THREADS) {
            const int r = idx / k, c = idx - r * k;
            Ls[r * LDL + c] = (c < r) ? L[(long long)r * ldl + c] : T(0);
        }
    };
    if (kp == 0) load_triangle();
    __syncthreads();
    // ---- collapse the k transpositions (k0+i <-> s_piv[i]) into row moves (same scheme as laswp.cu) ----
    if (tid < 2 * k) {
        int src = -1;
        if (tid < k) {
            src = k0 + tid;
        } else {
            const int u = tid - k;
            const int r = s_piv[u];
            if (r >= k0 + k) {  // a far row; take it once (first occurrence)
                bool dup = false;
                for (int v = 0; v < u; ++v) dup |= (s_piv[v] == r);
                if (!dup) src = r;
            }
        }
        if (src >= 0) {
            int cur = src;
            for (int i = 0; i < k; ++i) {
                const int ri = k0 + i, p = s_piv[i];
                if (cur == ri) cur = p;
                else if (cur == p) cur = ri;
            }
            if (cur < k0 + k) {
                s_topsrc[cur - k0] = src;            // lands in the top block (possibly unmoved)
            } else if (cur != src) {
                const int e = atomicAdd(&s_nfar, 1);  // leaves the top block
                s_farsrc[e] = src;
                s_fardst[e] = cur;
            }
        }
    }
    __syncthreads();
    const int nfar = s_nfar;
    // ---- gather: every source row segment read before anything is written ----
    for (int idx = tid; idx < k * COLS; idx += LT_THREADS) {
        const int i = idx / COLS, c = idx - i * COLS;
        if (c < cw) top[i * LDT + c] = A[(long long)s_topsrc[i] * lda + col0 + c];
    }
    for (int idx = tid; idx < nfar * COLS; idx += LT_THREADS) {
        const int e = idx / COLS, c = idx - e * COLS;
        if (c < cw) far[e * LDT + c] = A[(long long)s_farsrc[e] * lda + col0 + c];
    }
    __syncthreads();
    // ---- prefix: the kp rows directly above the block are already solved (an earlier call on this
    //      stream); the block's rows first take  -= L[block rows, prefix cols] * U[prefix rows] ----
    if (kp > 0) {
        const T* Lp = L - kp;  // same rows, kp columns to the left
        const int c = tid % COLS, rg = tid / COLS;
        constexpr int RG = LT_THREADS / COLS;       // row groups
        constexpr int RPT = (LT_KMAX + RG - 1) / RG;  // rows per thread
        for (int pc = 0; pc < kp; pc += PCH) {
            const int pw = (kp - pc) < PCH ? (kp - pc) : PCH;
            for (int idx = tid; idx < k * pw; idx += LT_THREADS) {
                const int r = idx / pw, t = idx - r * pw;
                Ls[r * LDL + t] = Lp[(long long)r * ldl + pc + t];
            }
            for (int idx = tid; idx < pw * COLS; idx += LT_THREADS) {
                const int t = idx / COLS, cc = idx - t * COLS;
                if (cc < cw) Ub[t * LDT + cc] = A[(long long)(k0 - kp + pc + t) * lda + col0 + cc];
            }
            __syncthreads();
            T v[RPT];
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
                const int i = rg + j * RG;
                v[j] = (i < k) ? top[i * LDT + c] : T(0);
            }
            // multipliers by 128-bit broadcast loads, PVEC prefix rows at a time (same FMAs in the same order)
            constexpr int PVEC = 16 / (int)sizeof(T);
            struct alignas(16) PV16 { T v[PVEC]; };
            for (int t = 0; t < pw; t += PVEC) {
                T ub[PVEC];
#pragma unroll
                for (int e = 0; e < PVEC; ++e) ub[e] = (t + e < pw) ? Ub[(t + e) * LDT + c] : T(0);
#pragma unroll
                for (int j = 0; j < RPT; ++j) {
                    const int i = rg + j * RG;
                    if (i < k) {
                        const PV16 l = *reinterpret_cast<const PV16*>(Ls + i * LDL + t);
#pragma unroll
                        for (int e = 0; e < PVEC; ++e)
                            if (t + e < pw) v[j] -= l.v[e] * ub[e];
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
                const int i = rg + j * RG;
                if (i < k) top[i * LDT + c] = v[j];
            }
            __syncthreads();
        }
        load_triangle();
        __syncthreads();
    }
    // ---- unit-lower solve of the top tile: row i -= L[i][kk] * row kk for kk ascending ----
    // Groups of 8 rows: one warp solves the group's own 8 x 8 triangle (lane = column, no block
    // barrier), then every thread applies the group to the rows below.  Each element sees exactly
    // the updates of the plain column sweep in the same order (bit-identical), with 2 block
    // barriers per 8 rows instead of one per row.
    constexpr int GR = 8;
    for (int g0 = 0; g0 < k; g0 += GR) {
        const int gk = (k - g0) < GR ? (k - g0) : GR;
        if (tid < COLS) {
            T x[GR];
#pragma unroll
            for (int r = 0; r < GR; ++r) x[r] = (r < gk) ? top[(g0 + r) * LDT + tid] : T(0);
#pragma unroll
            for (int r = 1; r < GR; ++r) {
                if (r < gk) {
#pragma unroll
                    for (int kk = 0; kk < r; ++kk) x[r] -= Ls[(g0 + r) * LDL + g0 + kk] * x[kk];
                    top[(g0 + r) * LDT + tid] = x[r];
                }
            }
        }
        __syncthreads();
        const int below = k - g0 - gk;  // rows under the group
        if (below > 0) {
            // register-tiled: the group's U entries of the thread's column stay in registers, every row's GR multipliers arrive
            // by 128-bit broadcast loads (one shared-memory load per 4 / 2 FMAs instead of two per FMA); same FMAs, same order
            constexpr int VEC = 16 / (int)sizeof(T);
            struct alignas(16) V16 { T v[VEC]; };
            const int c = tid % COLS, rg = tid / COLS;
            constexpr int RG = LT_THREADS / COLS;
            T ub[GR];
#pragma unroll
            for (int kk = 0; kk < GR; ++kk) ub[kk] = (kk < gk) ? top[(g0 + kk) * LDT + c] : T(0);
            for (int i = g0 + gk + rg; i < k; i += RG) {
                T v = top[i * LDT + c];
#pragma unroll
                for (int q = 0; q < GR / VEC; ++q) {
                    const V16 l = *reinterpret_cast<const V16*>(Ls + i * LDL + g0 + q * VEC);
#pragma unroll
                    for (int e = 0; e < VEC; ++e)
                        if (q * VEC + e < gk) v -= l.v[e] * ub[q * VEC + e];
                }
                top[i * LDT + c] = v;
            }
        }
        __syncthreads();
    }
    // ---- scatter ----
    for (int idx = tid; idx < k * COLS; idx += LT_THREADS) {
        const int i = idx / COLS, c = idx - i * COLS;
        if (c < cw) A[(long long)(k0 + i) * lda + col0 + c] = top[i * LDT + c];
    }
    for (int idx = tid; idx < nfar * COLS; idx += LT_THREADS) {
        const int e = idx / COLS, c = idx - e * COLS;
        if (c < cw) A[(long long)s_fardst[e] * lda + col0 + c] = far[e * LDT + c];
    }
}

template <class T, int KMAX>
int launch_laswp_trsm(int64_t ncols, T* d_a, int64_t lda, int64_t k0, int64_t k, const int32_t* d_ipiv, const T* d_l, int64_t ldl,
                      int64_t kp, cudaStream_t s) {
    auto kern = laswp_trsm_kernel<T, KMAX>;
    const size_t smem = LtCfg<T, KMAX>::smem_bytes;
    static bool configured = false;
    static uint64_t seen_epoch = 0;
    if (stale_for_context(seen_epoch)) configured = false;
    if (!configured) {
        LAIR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const unsigned grid = (unsigned)((ncols + LtCfg<T, KMAX>::COLS - 1) / LtCfg<T, KMAX>::COLS);
    ProfScope prof(kProfLaswp, s, 4.0 * (double)k * (double)ncols * sizeof(T));
    kern<<<grid, LT_THREADS, smem, s>>>(d_a, (long long)lda, (int)ncols, (int)k0, (int)k, d_ipiv, d_l, (long long)ldl, (int)kp);
    LAIR_LAUNCH_CHECK();
    return LAIR_B200_OK;
}

}  // namespace

// d_a points at row 0 of the matrix and the first of the `ncols` columns to update; the pivots
// ipiv[k0 .. k0+k) are absolute row indices; d_l is the k x k unit-lower block (leading dimension ldl).
// kp > 0: the kp rows above the block (k0-kp .. k0-1, same columns) are already solved and the block's
// rows take  -= L[k0.., k0-kp .. k0) * those rows  before their own triangle; d_l - kp must be that L block.
// A k x k step wider than 64 is then a chain of 64-row calls with growing prefix (blocked.cu).
// Returns LAIR_B200_ERR_UNSUPPORTED (no error text) when k > 128: callers fall back to laswp + trsm.
template <class T>
int laswp_trsm_dev(int64_t ncols, T* d_a, int64_t lda, int64_t k0, int64_t k, const int32_t* d_ipiv, const T* d_l, int64_t ldl,
                   cudaStream_t s, int64_t kp) {
    if (k > 128) return LAIR_B200_ERR_UNSUPPORTED;
    if (ncols <= 0 || k <= 0) return LAIR_B200_OK;
    LAIR_REQUIRE(ncols < (1ll << 31) && k0 + k < (1ll << 31), "laswp_trsm: dimension too large");
    LAIR_REQUIRE(kp >= 0 && kp <= k0, "laswp_trsm: bad prefix");
    if (k <= 64) return launch_laswp_trsm<T, 64>(ncols, d_a, lda, k0, k, d_ipiv, d_l, ldl, kp, s);
    return launch_laswp_trsm<T, 128>(ncols, d_a, lda, k0, k, d_ipiv, d_l, ldl, kp, s);
}

template int laswp_trsm_dev<float>(int64_t, float*, int64_t, int64_t, int64_t, const int32_t*, const float*, int64_t, cudaStream_t, int64_t);
template int laswp_trsm_dev<double>(int64_t, double*, int64_t, int64_t, int64_t, const int32_t*, const double*, int64_t, cudaStream_t, int64_t);

}  // namespace lair
