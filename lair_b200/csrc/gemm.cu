// f32 trailing-matrix update  C -= A * B  (row-major, alpha = -1, no conjugation); the f64
// tensor-core version lives in gemm_f64.cu.  Reference call site src/lapack/getrf.rs:289-296,
// routine src/blas/gemm.rs:6-32.
//
// Native FP32 FMA (exactly rounded products, meets the reference's backward error; a TF32x3
// tcgen05 path is the planned replacement): 128x128x16 CTA tiles staged by a 4-stage cp.async
// ring, 8x8 register tiles per thread.  Roofline: FP32-FMA-bound; flops = 2*M*N*K per launch.
#include "common.cuh"

namespace lair {
namespace {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 4, GEMM_THREADS = 256;

template <class T> struct Tile {
    static constexpr int VEC = 16 / sizeof(T);      // elements per 16-byte chunk
    static constexpr int LDA_S = BK + 4;            // f64: 20 doubles (bank = 8g+2t), f32: 20 floats
    static constexpr int LDB_S = BN + 4;            // f64: 132 doubles (bank = 8t+2g)
    static constexpr int A_ELEMS = BM * LDA_S;
    static constexpr int B_ELEMS = BK * LDB_S;
    static constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS;
    static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_ELEMS * sizeof(T);
};

// cp.async of `bytes` (0..16) valid source bytes into a 16-byte shared destination; the
// remainder is zero-filled, so out-of-range tile elements contribute nothing.
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gmem_src, int bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(bytes) : "memory");
}
template <int BYTES>
__device__ __forceinline__ void cp_async_small_zfill(void* smem_dst, const void* gmem_src, int bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;\n" ::"r"(s), "l"(gmem_src), "n"(BYTES), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// Stage loader shared by both precisions.  ALIGNED: every 16-byte chunk is 16-byte aligned
// in global memory (pointers and leading dimensions multiples of VEC).
template <class T, bool ALIGNED>
__device__ __forceinline__ void load_stage(T* __restrict__ sa, T* __restrict__ sb, const T* __restrict__ A, long long lda,
                                           const T* __restrict__ B, long long ldb, int M, int N, int K, int m0, int n0,
                                           int k0, int tid) {
    using TL = Tile<T>;
    constexpr int VEC = TL::VEC;
    // A tile: BM rows x BK cols
    constexpr int A_CPR = BK / VEC;  // chunks per row
    constexpr int A_CHUNKS = BM * A_CPR;
#pragma unroll
    for (int c = tid; c < A_CHUNKS; c += GEMM_THREADS) {
        int r = c / A_CPR, kc = (c % A_CPR) * VEC;
        int gr = m0 + r, gk = k0 + kc;
        int valid = (gr < M) ? (K - gk) : 0;
        valid = valid < 0 ? 0 : (valid > VEC ? VEC : valid);
        const T* src = A + (long long)(gr < M ? gr : 0) * lda + (valid > 0 ? gk : 0);
        T* dst = sa + r * TL::LDA_S + kc;
        if (ALIGNED) {
            cp_async16_zfill(dst, src, valid * (int)sizeof(T));
        } else {
#pragma unroll
            for (int e = 0; e < VEC; ++e)
                cp_async_small_zfill<sizeof(T)>(dst + e, (e < valid) ? src + e : src, (e < valid) ? (int)sizeof(T) : 0);
        }
    }
    // B tile: BK rows x BN cols
    constexpr int B_CPR = BN / VEC;
    constexpr int B_CHUNKS = BK * B_CPR;
#pragma unroll
    for (int c = tid; c < B_CHUNKS; c += GEMM_THREADS) {
        int r = c / B_CPR, nc = (c % B_CPR) * VEC;
        int gk = k0 + r, gn = n0 + nc;
        int valid = (gk < K) ? (N - gn) : 0;
        valid = valid < 0 ? 0 : (valid > VEC ? VEC : valid);
        const T* src = B + (long long)(gk < K ? gk : 0) * ldb + (valid > 0 ? gn : 0);
        T* dst = sb + r * TL::LDB_S + nc;
        if (ALIGNED) {
            cp_async16_zfill(dst, src, valid * (int)sizeof(T));
        } else {
#pragma unroll
            for (int e = 0; e < VEC; ++e)
                cp_async_small_zfill<sizeof(T)>(dst + e, (e < valid) ? src + e : src, (e < valid) ? (int)sizeof(T) : 0);
        }
    }
}

// ---- f32: FFMA, 8x8 register tiles -------------------------------------------------------------
template <bool ALIGNED>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
sgemm_minus_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ B, long long ldb,
                   float* __restrict__ C, long long ldc, int M, int N, int K, int tiles_m) {
    using TL = Tile<float>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* smem = reinterpret_cast<float*>(smem_raw);
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;  // rows ty + 16*i, cols tx*4 + 64*h + e
    const int tile = blockIdx.x;
    const int m0 = (tile % tiles_m) * BM, n0 = (tile / tiles_m) * BN;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int KT = (K + BK - 1) / BK;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) load_stage<float, ALIGNED>(smem + s * TL::STAGE_ELEMS, smem + s * TL::STAGE_ELEMS + TL::A_ELEMS, A, lda, B, ldb, M, N, K, m0, n0, s * BK, tid);
        cp_async_commit();
    }
    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nk = kt + STAGES - 1;
            if (nk < KT) {
                int slot = nk % STAGES;
                load_stage<float, ALIGNED>(smem + slot * TL::STAGE_ELEMS, smem + slot * TL::STAGE_ELEMS + TL::A_ELEMS, A, lda, B, ldb, M, N, K, m0, n0, nk * BK, tid);
            }
            cp_async_commit();
        }
        const float* sa = smem + (kt % STAGES) * TL::STAGE_ELEMS;
        const float* sb = sa + TL::A_ELEMS;
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float af[8], bf[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) af[i] = sa[(ty + 16 * i) * TL::LDA_S + kk];
            float4 b0 = *reinterpret_cast<const float4*>(sb + kk * TL::LDB_S + tx * 4);
            float4 b1 = *reinterpret_cast<const float4*>(sb + kk * TL::LDB_S + 64 + tx * 4);
            bf[0] = b0.x; bf[1] = b0.y; bf[2] = b0.z; bf[3] = b0.w;
            bf[4] = b1.x; bf[5] = b1.y; bf[6] = b1.z; bf[7] = b1.w;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(af[i], bf[j], acc[i][j]);
        }
    }
    cp_async_wait<0>();
    const bool vec_ok = ALIGNED && ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
    if (vec_ok && m0 + BM <= M && n0 + BN <= N) {
        // interior tile: all 16 loads of the C fragment in flight before the first use (a conditional
        // per element serialises one DRAM latency per load -- profiles/r1_dgemm_ncu.md)
        float* cbase = C + (long long)(m0 + ty) * ldc + n0 + tx * 4;
        float4 cv[8][2];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h) cv[i][h] = *reinterpret_cast<const float4*>(cbase + (long long)(16 * i) * ldc + h * 64);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float4 v = cv[i][h];
                v.x -= acc[i][h * 4 + 0];
                v.y -= acc[i][h * 4 + 1];
                v.z -= acc[i][h * 4 + 2];
                v.w -= acc[i][h * 4 + 3];
                *reinterpret_cast<float4*>(cbase + (long long)(16 * i) * ldc + h * 64) = v;
            }
        return;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int row = m0 + ty + 16 * i;
        if (row >= M) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int col = n0 + h * 64 + tx * 4;
            float* p = C + (long long)row * ldc + col;
            if (vec_ok && col + 3 < N) {
                float4 v = *reinterpret_cast<float4*>(p);
                v.x -= acc[i][h * 4 + 0];
                v.y -= acc[i][h * 4 + 1];
                v.z -= acc[i][h * 4 + 2];
                v.w -= acc[i][h * 4 + 3];
                *reinterpret_cast<float4*>(p) = v;
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (col + e < N) p[e] -= acc[i][h * 4 + e];
            }
        }
    }
}

template <class T> struct GemmKernel;
template <> struct GemmKernel<float> {
    template <bool AL> static auto get() { return sgemm_minus_kernel<AL>; }
};

template <class T, bool AL>
int launch_gemm(int64_t m, int64_t n, int64_t k, const T* d_a, int64_t lda, const T* d_b, int64_t ldb, T* d_c, int64_t ldc,
                cudaStream_t s) {
    auto kern = GemmKernel<T>::template get<AL>();
    static bool configured = false;
    static uint64_t seen_epoch = 0;
    if (stale_for_context(seen_epoch)) configured = false;
    if (!configured) {
        LAIR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tile<T>::SMEM_BYTES));
        configured = true;
    }
    int64_t tiles_m = (m + BM - 1) / BM, tiles_n = (n + BN - 1) / BN;
    int64_t tiles = tiles_m * tiles_n;
    LAIR_REQUIRE(tiles < (1ll << 31), "gemm: too many tiles");
    ProfScope prof(kProfGemm, s, 2.0 * (double)m * (double)n * (double)k);
    kern<<<(unsigned)tiles, GEMM_THREADS, Tile<T>::SMEM_BYTES, s>>>(d_a, (long long)lda, d_b, (long long)ldb, d_c, (long long)ldc, (int)m,
                                                                  (int)n, (int)k, (int)tiles_m);
    LAIR_LAUNCH_CHECK();
    return LAIR_B200_OK;
}

}  // namespace

template <class T>
int gemm_minus_dev(int64_t m, int64_t n, int64_t k, const T* d_a, int64_t lda, const T* d_b, int64_t ldb, T* d_c, int64_t ldc,
                   cudaStream_t s) {
    LAIR_REQUIRE(m >= 0 && n >= 0 && k >= 0, "gemm: negative dimension");
    LAIR_REQUIRE(m < (1ll << 31) && n < (1ll << 31) && k < (1ll << 31), "gemm: dimension too large");
    if (m == 0 || n == 0 || k == 0) return LAIR_B200_OK;
    LAIR_REQUIRE(lda >= k && ldb >= n && ldc >= n, "gemm: leading dimension too small");
    if constexpr (sizeof(T) == 4) {
        // large updates go to the tensor cores (3xTF32 split, gemm_tf32.cu); one workspace per stream of the sweep
        if (ctx().opt.sgemm_tf32 && sgemm_tf32x3_supported(m, n, k, d_c, ldc)) {
            const int slot = s == ctx().aux_stream ? Context::kWorkTf32Aux : (s == ctx().stream ? Context::kWorkTf32Main : Context::kWorkTf32Other);
            return sgemm_tf32x3_minus_dev(m, n, k, d_a, lda, d_b, ldb, d_c, ldc, slot, s);
        }
    }
    constexpr int VEC = 16 / sizeof(T);
    const bool aligned = (lda % VEC == 0) && (ldb % VEC == 0) && (reinterpret_cast<uintptr_t>(d_a) % 16 == 0) &&
                         (reinterpret_cast<uintptr_t>(d_b) % 16 == 0);
    if (aligned) return launch_gemm<T, true>(m, n, k, d_a, lda, d_b, ldb, d_c, ldc, s);
    return launch_gemm<T, false>(m, n, k, d_a, lda, d_b, ldb, d_c, ldc, s);
}

template int gemm_minus_dev<float>(int64_t, int64_t, int64_t, const float*, int64_t, const float*, int64_t, float*, int64_t, cudaStream_t);

}  // namespace lair
