// f64 trailing-matrix update  C -= A * B  (row-major): the one true contraction of the LU
// path (reference call site src/lapack/getrf.rs:289-296, routine src/blas/gemm.rs:6-32).
//
// FP64 tensor cores: tcgen05.mma has no f64 kind on sm_100a, so the tensor path is
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4 -- the only native DMMA shape on this chip; measured
// issue-rate peak 37.0 TFLOP/s, profiles/r1_microbench_fp64_peak.jsonl) with register
// accumulators.  Operands are staged global -> shared with a cp.async (LDGSTS) ring whose
// row pitches (BK+4, BN+4 doubles) make every 64-bit fragment load bank-conflict free.
//
// Tile configurations (template): CTA tile = (WARPS_M*WTM*8) x (WARPS_N*WTN*8), BK = 16.
//   big    128 x 64, 4 warps of 64x32: two CTAs per SM, so one CTA's prologue/epilogue hides
//          under the other's DMMA main loop;
//   skinny  64 x 32, 2 warps of 32x32: panel-internal and multi-RHS updates with N <= 64,
//          where the big tile would leave most SMs idle.
// The accumulators start at -C (loaded before the main loop, overlapping the first operand
// loads), so the epilogue is a pure store of -(acc) = C - A*B.
// Roofline: tensor-bound; algorithmic flops = 2*M*N*K per launch.
#include "common.cuh"

namespace lair {
namespace {

constexpr int BK = 16;

__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gmem_src, int bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async8_zfill(void* smem_dst, const void* gmem_src, int bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <int WARPS_M, int WARPS_N, int WTM, int WTN, int STAGES_>
struct Cfg {
    static constexpr int THREADS = WARPS_M * WARPS_N * 32;
    static constexpr int BM = WARPS_M * WTM * 8, BN = WARPS_N * WTN * 8;
    static constexpr int STAGES = STAGES_;
    static constexpr int LDA_S = BK + 4;  // doubles: fragment bank = 8g + 2t
    static constexpr int LDB_S = BN + 4;  // doubles: fragment bank = 8t + 2g
    static constexpr int A_ELEMS = BM * LDA_S, B_ELEMS = BK * LDB_S;
    static constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS;
    static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_ELEMS * sizeof(double);
};

template <class C, bool ALIGNED>
__device__ __forceinline__ void load_stage(double* __restrict__ sa, double* __restrict__ sb, const double* __restrict__ A,
                                           long long lda, const double* __restrict__ B, long long ldb, int M, int N, int K,
                                           int m0, int n0, int k0, int tid) {
    constexpr int A_CPR = BK / 2, A_CHUNKS = C::BM * A_CPR;
#pragma unroll
    for (int c = tid; c < A_CHUNKS; c += C::THREADS) {
        const int r = c / A_CPR, kc = (c % A_CPR) * 2;
        const int gr = m0 + r, gk = k0 + kc;
        int valid = (gr < M) ? (K - gk) : 0;
        valid = valid < 0 ? 0 : (valid > 2 ? 2 : valid);
        const double* src = A + (long long)(gr < M ? gr : 0) * lda + (valid > 0 ? gk : 0);
        double* dst = sa + r * C::LDA_S + kc;
        if (ALIGNED) {
            cp_async16_zfill(dst, src, valid * 8);
        } else {
            cp_async8_zfill(dst, src, valid > 0 ? 8 : 0);
            cp_async8_zfill(dst + 1, valid > 1 ? src + 1 : src, valid > 1 ? 8 : 0);
        }
    }
    constexpr int B_CPR = C::BN / 2, B_CHUNKS = BK * B_CPR;
#pragma unroll
    for (int c = tid; c < B_CHUNKS; c += C::THREADS) {
        const int r = c / B_CPR, nc = (c % B_CPR) * 2;
        const int gk = k0 + r, gn = n0 + nc;
        int valid = (gk < K) ? (N - gn) : 0;
        valid = valid < 0 ? 0 : (valid > 2 ? 2 : valid);
        const double* src = B + (long long)(gk < K ? gk : 0) * ldb + (valid > 0 ? gn : 0);
        double* dst = sb + r * C::LDB_S + nc;
        if (ALIGNED) {
            cp_async16_zfill(dst, src, valid * 8);
        } else {
            cp_async8_zfill(dst, src, valid > 0 ? 8 : 0);
            cp_async8_zfill(dst + 1, valid > 1 ? src + 1 : src, valid > 1 ? 8 : 0);
        }
    }
}

// Interior tiles with 16-byte-aligned operands: per-thread source pointers are fixed at kernel
// start and only advance by BK (A) / BK rows (B) per k-tile -- no bounds logic in the loop.
template <class C>
__device__ __forceinline__ void load_stage_fast(double* __restrict__ sa, double* __restrict__ sb, const double* __restrict__ a_src,
                                                long long lda, const double* __restrict__ b_src, long long ldb, int tid) {
    constexpr int A_CPR = BK / 2, A_ITERS = C::BM * A_CPR / C::THREADS, A_ROWS_PER_ITER = C::THREADS / A_CPR;
    static_assert(C::BM * A_CPR % C::THREADS == 0 && C::THREADS % A_CPR == 0, "A tile must divide evenly");
    double* da = sa + (tid / A_CPR) * C::LDA_S + (tid % A_CPR) * 2;
#pragma unroll
    for (int i = 0; i < A_ITERS; ++i)
        cp_async16_zfill(da + i * A_ROWS_PER_ITER * C::LDA_S, a_src + (long long)(i * A_ROWS_PER_ITER) * lda, 16);
    constexpr int B_CPR = C::BN / 2;
    if constexpr (C::THREADS % B_CPR == 0) {
        constexpr int B_ITERS = BK * B_CPR / C::THREADS, B_ROWS_PER_ITER = C::THREADS / B_CPR;
        static_assert(BK * B_CPR % C::THREADS == 0, "B tile must divide evenly");
        double* db = sb + (tid / B_CPR) * C::LDB_S + (tid % B_CPR) * 2;
#pragma unroll
        for (int i = 0; i < B_ITERS; ++i)
            cp_async16_zfill(db + i * B_ROWS_PER_ITER * C::LDB_S, b_src + (long long)(i * B_ROWS_PER_ITER) * ldb, 16);
    } else {
        // fewer threads than chunks per row: each thread walks columns within a row
        constexpr int B_CHUNKS = BK * B_CPR;
#pragma unroll
        for (int c = tid; c < B_CHUNKS; c += C::THREADS) {
            const int r = c / B_CPR, nc = (c % B_CPR) * 2;
            cp_async16_zfill(sb + r * C::LDB_S + nc, b_src + (long long)r * ldb + nc, 16);
        }
    }
}

template <class C, int WARPS_M, int WARPS_N, int WTM, int WTN, int MINB, bool ALIGNED>
__global__ void __launch_bounds__(C::THREADS, MINB)
dgemm_minus_kernel(const double* __restrict__ A, long long lda, const double* __restrict__ B, long long ldb,
                   double* __restrict__ Cm, long long ldc, int M, int N, int K, int tiles_m, int tiles_n, int raster) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* smem = reinterpret_cast<double*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp / WARPS_N, wn = warp % WARPS_N;
    // CTA order: strips of `raster` tile columns, row by row inside a strip.  The CTAs resident at any time then cover
    // (resident / raster) tile rows x raster tile columns, so an A (L21) tile is read from DRAM once per STRIP instead of
    // once per tile column -- at M = 65 536, K = 256 the A panel is 134 MB, larger than the L2, and walking down M one
    // tile column at a time (raster = 1) streamed it from DRAM for every one of the N / BN tile columns.
    const int tile = blockIdx.x;
    const int per_strip = tiles_m * raster;
    const int strip = tile / per_strip, rem = tile - strip * per_strip;
    const int sw = min(raster, tiles_n - strip * raster);  // the last strip may be narrower
    const int m0 = (rem / sw) * C::BM, n0 = (strip * raster + rem % sw) * C::BN;
    const int KT = (K + BK - 1) / BK;

    // interior tile with aligned operands: fixed per-thread source pointers, no bounds logic
    const bool fast = ALIGNED && (m0 + C::BM <= M) && (n0 + C::BN <= N) && (K % BK == 0);
    constexpr int A_CPR = BK / 2, B_CPR = C::BN / 2;
    const double* a_thr = A + (long long)(m0 + tid / A_CPR) * lda + (tid % A_CPR) * 2;
    const double* b_thr = (C::THREADS % B_CPR == 0) ? B + (long long)(tid / B_CPR) * ldb + n0 + (tid % B_CPR) * 2 : B + n0;
    auto issue_stage = [&](int kt_load) {
        const int slot = kt_load % C::STAGES;
        double* sa_ = smem + slot * C::STAGE_ELEMS;
        double* sb_ = sa_ + C::A_ELEMS;
        if (fast)
            load_stage_fast<C>(sa_, sb_, a_thr + (long long)kt_load * BK, lda, b_thr + (long long)kt_load * BK * ldb, ldb, tid);
        else
            load_stage<C, ALIGNED>(sa_, sb_, A, lda, B, ldb, M, N, K, m0, n0, kt_load * BK, tid);
    };

    // operand pipeline first, so the C loads below overlap it
#pragma unroll
    for (int s = 0; s < C::STAGES - 1; ++s) {
        if (s < KT) issue_stage(s);
        cp_async_commit();
    }

    // accumulators start at -C: each thread owns (row g, cols 2t, 2t+1) of every 8x8 tile
    double acc[WTM][WTN][2];
    const bool vec_ok = ALIGNED && ((ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(Cm) & 15) == 0);
    // interior tiles (the common case) take a branch-free path so all C loads are in flight at
    // once; a conditional per element would serialise one DRAM latency per load
    const bool interior = vec_ok && (m0 + C::BM <= M) && (n0 + C::BN <= N);
    if (interior) {
        const double* cbase = Cm + (long long)(m0 + wm * WTM * 8 + g) * ldc + n0 + wn * WTN * 8 + 2 * t;
        double2 cv[WTM][WTN];
#pragma unroll
        for (int i = 0; i < WTM; ++i)
#pragma unroll
            for (int j = 0; j < WTN; ++j) cv[i][j] = *reinterpret_cast<const double2*>(cbase + (long long)(i * 8) * ldc + j * 8);
#pragma unroll
        for (int i = 0; i < WTM; ++i)
#pragma unroll
            for (int j = 0; j < WTN; ++j) {
                acc[i][j][0] = -cv[i][j].x;
                acc[i][j][1] = -cv[i][j].y;
            }
    } else {
#pragma unroll
        for (int i = 0; i < WTM; ++i) {
            const int row = m0 + (wm * WTM + i) * 8 + g;
#pragma unroll
            for (int j = 0; j < WTN; ++j) {
                const int col = n0 + (wn * WTN + j) * 8 + 2 * t;
                double c0 = 0.0, c1 = 0.0;
                if (row < M) {
                    const double* p = Cm + (long long)row * ldc + col;
                    if (col < N) c0 = p[0];
                    if (col + 1 < N) c1 = p[1];
                }
                acc[i][j][0] = -c0;
                acc[i][j][1] = -c1;
            }
        }
    }

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<C::STAGES - 2>();
        __syncthreads();
        {
            const int nk = kt + C::STAGES - 1;
            if (nk < KT) issue_stage(nk);
            cp_async_commit();
        }
        const double* sa = smem + (kt % C::STAGES) * C::STAGE_ELEMS;
        const double* sb = sa + C::A_ELEMS;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double af[WTM], bf[WTN];
#pragma unroll
            for (int i = 0; i < WTM; ++i) af[i] = sa[((wm * WTM + i) * 8 + g) * C::LDA_S + kk * 4 + t];
#pragma unroll
            for (int j = 0; j < WTN; ++j) bf[j] = sb[(kk * 4 + t) * C::LDB_S + (wn * WTN + j) * 8 + g];
#pragma unroll
            for (int i = 0; i < WTM; ++i)
#pragma unroll
                for (int j = 0; j < WTN; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: C = -(acc)
    if (interior) {
        double* cbase = Cm + (long long)(m0 + wm * WTM * 8 + g) * ldc + n0 + wn * WTN * 8 + 2 * t;
#pragma unroll
        for (int i = 0; i < WTM; ++i)
#pragma unroll
            for (int j = 0; j < WTN; ++j)
                *reinterpret_cast<double2*>(cbase + (long long)(i * 8) * ldc + j * 8) = make_double2(-acc[i][j][0], -acc[i][j][1]);
        return;
    }
#pragma unroll
    for (int i = 0; i < WTM; ++i) {
        const int row = m0 + (wm * WTM + i) * 8 + g;
        if (row >= M) continue;
#pragma unroll
        for (int j = 0; j < WTN; ++j) {
            const int col = n0 + (wn * WTN + j) * 8 + 2 * t;
            double* p = Cm + (long long)row * ldc + col;
            if (vec_ok && col + 1 < N) {
                *reinterpret_cast<double2*>(p) = make_double2(-acc[i][j][0], -acc[i][j][1]);
            } else {
                if (col < N) p[0] = -acc[i][j][0];
                if (col + 1 < N) p[1] = -acc[i][j][1];
            }
        }
    }
}

template <int WARPS_M, int WARPS_N, int WTM, int WTN, int STAGES_, int MINB, bool AL>
int launch(int64_t m, int64_t n, int64_t k, const double* d_a, int64_t lda, const double* d_b, int64_t ldb, double* d_c,
           int64_t ldc, cudaStream_t s) {
    using C = Cfg<WARPS_M, WARPS_N, WTM, WTN, STAGES_>;
    auto kern = dgemm_minus_kernel<C, WARPS_M, WARPS_N, WTM, WTN, MINB, AL>;
    static bool configured = false;
    static uint64_t seen_epoch = 0;
    if (stale_for_context(seen_epoch)) configured = false;
    if (!configured) {
        LAIR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM_BYTES));
        configured = true;
    }
    const int64_t tiles_m = (m + C::BM - 1) / C::BM, tiles_n = (n + C::BN - 1) / C::BN;
    const int64_t tiles = tiles_m * tiles_n;
    LAIR_REQUIRE(tiles < (1ll << 31), "gemm: too many tiles");
    ProfScope prof(kProfGemm, s, 2.0 * (double)m * (double)n * (double)k);
    int64_t raster = ctx().opt.gemm_raster;
    if (raster > tiles_n) raster = tiles_n;
    kern<<<(unsigned)tiles, C::THREADS, C::SMEM_BYTES, s>>>(d_a, (long long)lda, d_b, (long long)ldb, d_c, (long long)ldc, (int)m, (int)n,
                                                           (int)k, (int)tiles_m, (int)tiles_n, (int)raster);
    LAIR_LAUNCH_CHECK();
    return LAIR_B200_OK;
}

template <bool AL>
int dispatch(int64_t m, int64_t n, int64_t k, const double* d_a, int64_t lda, const double* d_b, int64_t ldb, double* d_c,
             int64_t ldc, cudaStream_t s) {
    const int64_t cfg = ctx().opt.gemm_cfg;
    // tiles of the big configuration; if they cannot fill the SMs twice over, go skinny
    const int64_t big_tiles = ((m + 127) / 128) * ((n + 63) / 64);
    const bool skinny = (cfg == 2) || (cfg == 0 && (n <= 32 || big_tiles < 2 * (int64_t)ctx().sm_count));
    if (skinny) return launch<2, 1, 4, 4, 3, 4, AL>(m, n, k, d_a, lda, d_b, ldb, d_c, ldc, s);   // 64 x 32, 64 threads
    if (cfg == 3) return launch<2, 4, 8, 4, 4, 1, AL>(m, n, k, d_a, lda, d_b, ldb, d_c, ldc, s); // 128 x 128, 256 threads, 1 CTA/SM
    return launch<2, 2, 8, 4, 3, 2, AL>(m, n, k, d_a, lda, d_b, ldb, d_c, ldc, s);               // 128 x 64, 128 threads, 2 CTAs/SM
}

}  // namespace

template <>
int gemm_minus_dev<double>(int64_t m, int64_t n, int64_t k, const double* d_a, int64_t lda, const double* d_b, int64_t ldb,
                           double* d_c, int64_t ldc, cudaStream_t s) {
    LAIR_REQUIRE(m >= 0 && n >= 0 && k >= 0, "gemm: negative dimension");
    LAIR_REQUIRE(m < (1ll << 31) && n < (1ll << 31) && k < (1ll << 31), "gemm: dimension too large");
    if (m == 0 || n == 0 || k == 0) return LAIR_B200_OK;
    LAIR_REQUIRE(lda >= k && ldb >= n && ldc >= n, "gemm: leading dimension too small");
    const bool aligned = (lda % 2 == 0) && (ldb % 2 == 0) && (reinterpret_cast<uintptr_t>(d_a) % 16 == 0) &&
                         (reinterpret_cast<uintptr_t>(d_b) % 16 == 0);
    if (aligned) return dispatch<true>(m, n, k, d_a, lda, d_b, ldb, d_c, ldc, s);
    return dispatch<false>(m, n, k, d_a, lda, d_b, ldb, d_c, ldc, s);
}

}  // namespace lair
