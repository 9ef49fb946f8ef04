// Unit-lower triangular solve of one block step's U12 rows in ONE launch, k <= 256:
//     B[0..k, cols]  <-  L11^-1 * B[0..k, cols]          (src/blas/trsm.rs:6-22; getrf.rs:278-283)
// for the wide column ranges of the trailing update (the rows have had their interchanges already:
// laswp.cu).  Replaces, on those ranges, the chain of 64-row fused launches of laswp_trsm.cu, whose
// prefix update read both operands of every FMA from shared memory (measured: the "laswp" family took
// 30 ms of a 53 ms sgetrf at n = 16 384, 12.6 ms of a 24.4 ms dgetrf at n = 8 192, all of it serial
// with the GEMM on the main stream; profiles/r2t_probe_families.jsonl).
//
// One CTA per strip of CW columns, the strip's k rows resident in shared memory; L11 streams through
// shared memory in blocks of LB columns.  Per block: (1) the LB x LB triangle, 8 rows at a time (one
// warp per 32 columns solves the 8 x 8 triangle with lane = column, then all threads eliminate the
// group from the block's remaining rows); (2) the rows below the block take  -= L[rows, block] * U[block]
// as a register-tiled update: every thread keeps the block's U entries of its column(s) in registers
// (LB values per column) and walks its rows, fetching each row's LB multipliers with 128-bit broadcast
// loads -- one shared-memory load per 4 (f32) / 2 (f64) FMAs instead of two per FMA.
// Every element still sees  x -= l * u  as one FMA per eliminated row in ascending order, so the result
// is bit-identical to the plain column sweep and to laswp_trsm.cu (tests/test_gpu_parity.py).
#include "common.cuh"

namespace lair {
namespace {

constexpr int TS_THREADS = 256;
constexpr int TS_KMAX = 256;
constexpr int TS_GR = 8;  // rows per triangle group

template <class T> struct TsCfg;
template <> struct TsCfg<float> {
    static constexpr int CPT = 2;    // columns per thread
    static constexpr int LB = 32;    // L columns per block
    static constexpr int LP = 36;    // L tile pitch (16-byte aligned rows)
};
template <> struct TsCfg<double> {
    static constexpr int CPT = 1;
    static constexpr int LB = 16;
    static constexpr int LP = 18;
};

template <class T>
__global__ void __launch_bounds__(TS_THREADS, 2)
trsm_strip_kernel(const T* __restrict__ L, long long ldl, T* __restrict__ B, long long ldb, int k, int ncols) {
    using C = TsCfg<T>;
    constexpr int CPT = C::CPT, LB = C::LB, LP = C::LP;
    constexpr int CW = 32 * CPT, LDT = CW + 1;
    constexpr int VEC = 16 / sizeof(T);
    constexpr int RG = TS_THREADS / 32;  // row groups of the update phase = warps
    constexpr int RCH = 8;               // rows a thread updates per pass
    struct alignas(16) V16 { T v[VEC]; };
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* top = reinterpret_cast<T*>(smem_raw);                       // [k][LDT]
    T* Ls = top + ((size_t)k * LDT + VEC - 1) / VEC * VEC;         // [k - r0][LP]: rows r0.. of the current L block

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int col0 = blockIdx.x * CW;
    const int cw = (ncols - col0) < CW ? (ncols - col0) : CW;

    for (int idx = tid; idx < k * CW; idx += TS_THREADS) {
        const int i = idx / CW, c = idx - i * CW;
        top[i * LDT + c] = (c < cw) ? B[(long long)i * ldb + col0 + c] : T(0);
    }
    for (int r0 = 0; r0 < k; r0 += LB) {
        const int bk = (k - r0) < LB ? (k - r0) : LB;  // rows (= L columns) of this block
        __syncthreads();                               // top loaded / previous block's updates done, Ls free
        for (int idx = tid; idx < (k - r0) * LB; idx += TS_THREADS) {
            const int i = idx / LB, t = idx - i * LB;
            Ls[i * LP + t] = (t < bk) ? L[(long long)(r0 + i) * ldl + r0 + t] : T(0);
        }
        __syncthreads();
        // ---- (1) the block's own triangle, TS_GR rows at a time ----
        for (int g0 = 0; g0 < bk; g0 += TS_GR) {
            const int gk = (bk - g0) < TS_GR ? (bk - g0) : TS_GR;
            if (warp < CPT) {
                const int c = warp * 32 + lane;
                T x[TS_GR];
#pragma unroll
                for (int r = 0; r < TS_GR; ++r) x[r] = (r < gk) ? top[(r0 + g0 + r) * LDT + c] : T(0);
#pragma unroll
                for (int r = 1; r < TS_GR; ++r) {
                    if (r < gk) {
#pragma unroll
                        for (int kk = 0; kk < r; ++kk) x[r] -= Ls[(g0 + r) * LP + g0 + kk] * x[kk];
                        top[(r0 + g0 + r) * LDT + c] = x[r];
                    }
                }
            }
            __syncthreads();
            const int below = bk - g0 - gk;  // rows of the block under the group
            for (int idx = tid; idx < below * CW; idx += TS_THREADS) {
                const int i = g0 + gk + idx / CW, c = idx % CW;
                T v = top[(r0 + i) * LDT + c];
                for (int kk = 0; kk < gk; ++kk) v -= Ls[i * LP + g0 + kk] * top[(r0 + g0 + kk) * LDT + c];
                top[(r0 + i) * LDT + c] = v;
            }
            if (below > 0) __syncthreads();
        }
        // ---- (2) rows below the block: -= L[rows, block] * U[block], U of the thread's columns in registers ----
        const int nbelow = k - r0 - bk;
        if (nbelow > 0) {
            __syncthreads();  // (the last group's rows are final)
            T ub[LB][CPT];
#pragma unroll
            for (int t = 0; t < LB; ++t) {
#pragma unroll
                for (int p = 0; p < CPT; ++p) ub[t][p] = top[(r0 + t) * LDT + lane + 32 * p];  // rows >= bk of a ragged block: their L entries are zero
            }
            for (int ib = warp; ib < nbelow; ib += RG * RCH) {
                T v[RCH][CPT];
#pragma unroll
                for (int j = 0; j < RCH; ++j) {
                    const int i = ib + j * RG;
#pragma unroll
                    for (int p = 0; p < CPT; ++p) v[j][p] = (i < nbelow) ? top[(r0 + bk + i) * LDT + lane + 32 * p] : T(0);
                }
#pragma unroll
                for (int t = 0; t < LB; t += VEC) {
#pragma unroll
                    for (int j = 0; j < RCH; ++j) {
                        const int i = ib + j * RG;
                        const int il = (i < nbelow) ? (bk + i) : bk;  // any valid row for the idle slots
                        const V16 l = *reinterpret_cast<const V16*>(Ls + il * LP + t);
#pragma unroll
                        for (int e = 0; e < VEC; ++e) {
#pragma unroll
                            for (int p = 0; p < CPT; ++p) v[j][p] -= l.v[e] * ub[t + e][p];
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < RCH; ++j) {
                    const int i = ib + j * RG;
                    if (i < nbelow) {
#pragma unroll
                        for (int p = 0; p < CPT; ++p) top[(r0 + bk + i) * LDT + lane + 32 * p] = v[j][p];
                    }
                }
            }
        }
    }
    __syncthreads();
    for (int idx = tid; idx < k * CW; idx += TS_THREADS) {
        const int i = idx / CW, c = idx - i * CW;
        if (i > 0 && c < cw) B[(long long)i * ldb + col0 + c] = top[i * LDT + c];  // row 0 is unchanged by a unit-lower solve
    }
}

template <class T>
size_t ts_smem_bytes(int64_t k) {
    using C = TsCfg<T>;
    constexpr int VEC = 16 / sizeof(T);
    const size_t top = ((size_t)k * (32 * C::CPT + 1) + VEC - 1) / VEC * VEC;
    return (top + (size_t)k * C::LP) * sizeof(T);
}

}  // namespace

// B (k x ncols, row-major) <- L^-1 B with L the k x k unit-lower block at d_l.  Returns LAIR_B200_ERR_UNSUPPORTED
// (no error text) for k > 256: callers fall back to the recursive solve (trsm.cu).
template <class T>
int trsm_strip_dev(int64_t k, int64_t ncols, const T* d_l, int64_t ldl, T* d_b, int64_t ldb, cudaStream_t s) {
    if (k > TS_KMAX) return LAIR_B200_ERR_UNSUPPORTED;
    if (k <= 1 || ncols <= 0) return LAIR_B200_OK;
    LAIR_REQUIRE(ncols < (1ll << 31), "trsm_strip: dimension too large");
    auto kern = trsm_strip_kernel<T>;
    static bool configured = false;
    static uint64_t seen_epoch = 0;
    if (stale_for_context(seen_epoch)) configured = false;
    if (!configured) {
        LAIR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ts_smem_bytes<T>(TS_KMAX)));
        configured = true;
    }
    constexpr int CW = 32 * TsCfg<T>::CPT;
    const unsigned grid = (unsigned)((ncols + CW - 1) / CW);
    ProfScope prof(kProfTrsm, s, (double)k * (double)k * (double)ncols);
    kern<<<grid, TS_THREADS, ts_smem_bytes<T>(k), s>>>(d_l, (long long)ldl, d_b, (long long)ldb, (int)k, (int)ncols);
    LAIR_LAUNCH_CHECK();
    return LAIR_B200_OK;
}

template int trsm_strip_dev<float>(int64_t, int64_t, const float*, int64_t, float*, int64_t, cudaStream_t);
template int trsm_strip_dev<double>(int64_t, int64_t, const double*, int64_t, double*, int64_t, cudaStream_t);

}  // namespace lair
