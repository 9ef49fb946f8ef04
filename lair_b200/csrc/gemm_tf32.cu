// f32 trailing-matrix update  C -= A * B  on the 5th-generation tensor cores: tcgen05.mma.kind::tf32 with the
// operands split in two TF32 terms each (3 products: hi*hi + hi*lo + lo*hi, "3xTF32"), operands staged by TMA into
// 128B-swizzled shared memory, accumulator in TMEM, read back with tcgen05.ld for the epilogue.
// Reference: the contraction of src/blas/gemm.rs:6-32 at its LU call site src/lapack/getrf.rs:289-296.
//
// Why a split: TF32 keeps 11 significant bits.  a = a_hi + a_lo (+ 2^-22 a) with a_hi = tf32(a), a_lo = tf32(a - a_hi),
// so a*b = a_hi*b_hi + a_hi*b_lo + a_lo*b_hi up to ~2^-21 relative -- the products themselves are exact in the f32
// accumulator.  The result is NOT the reference's bit pattern (neither is the FFMA kernel's, which sums in a different
// order); the f32 blocked path is held to the backward-error bar (<= 10x the oracle's), not to bit equality.
//
// Shape of the kernel (one 128 x 128 output tile per CTA, 192 threads):
//   warp 0, one elected lane : TMA producer -- for each of the K/16 k-blocks, the 128x16 boxes of A_hi, A_lo, B^T_hi, B^T_lo
//                              (K-major, 64 B rows, SWIZZLE_64B) into a 3-stage ring, completion on an mbarrier: the four
//                              tiles travel ONCE and feed all three products (the first version loaded a pair per product:
//                              393 KB of operands per output tile at K = 128 made the kernel L2-bandwidth-bound);
//   warp 1, one elected lane : MMA issuer -- 3 x 2 tcgen05.mma (M 128, N 128, K 8) per k-block on shared-memory descriptors,
//                              tcgen05.commit frees the stage; the last commit signals the epilogue;
//   warps 2-5               : epilogue -- TMEM -> registers (tcgen05.ld 32x32b.x32) -> shared memory -> C -= acc in full
//                              128-byte row segments.
// Two pre-passes build the split operands: A (m x k) -> A_hi, A_lo; B (k x n) -> B^T_hi, B^T_lo (n x k), so both MMA
// operands are K-major.  Their traffic is O((m + n) k) against the O(m n) of C.
// With the flops on tensor cores the update is HBM-bound on C: algorithmic bytes = 2 * m * n * 4 per launch.
#include <cuda.h>

#include "common.cuh"

namespace lair {
namespace {

constexpr int BM = 128, BN = 128, BK = 16;  // BK tf32 = 64 bytes = one SWIZZLE_64B atom row
constexpr int kStages = 3;  // 3 x 32 KB of operand tiles per CTA: two CTAs per SM, one CTA's epilogue under the other's main loop
constexpr int kTileBytes = BM * BK * 4;      // 8 KB per operand tile; a stage holds A_hi, A_lo, B_hi, B_lo of one k-block
constexpr int kStageBytes = 4 * kTileBytes;
constexpr int kThreads = 192;
constexpr uint32_t kTmemCols = 256;  // columns 0-127: hi*hi; 128-255: the two small products, summed separately and added once

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a pipeline bug traps (the launch fails loudly) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (int spin = 0; spin < (1 << 24); ++spin) {
        uint32_t ok;
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
    }
    asm volatile("trap;");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major operand tile, 64-byte rows, SWIZZLE_64B: 8-row groups 512 bytes apart (SBO), LBO unused, descriptor version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
// kind::tf32, D = f32, A and B K-major, M = 128, N = 128
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

// one warp reads 32 TMEM lanes x 32 consecutive columns: lane i receives its row's 32 values
__device__ __forceinline__ void tmem_ld32(uint32_t (&v)[32], uint32_t taddr) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                   "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                   "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}

struct __align__(8) Bars {
    unsigned long long full[kStages], empty[kStages], tmem_full;
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(kThreads)
sgemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_alo,
                    const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo,
                    float* __restrict__ C, long long ldc, int M, int N, int K, int tiles_m) {
    extern __shared__ unsigned char smem_raw[];
    // tiles need 1024-byte alignment (swizzle atom = 8 rows x 128 bytes)
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* gen = smem_raw + (base - raw);
    Bars* bars = reinterpret_cast<Bars*>(gen + kStages * kStageBytes);
    const uint32_t sa0 = base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = (blockIdx.x % tiles_m) * BM, n0 = (blockIdx.x / tiles_m) * BN;
    const int nkb = K / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(smem_u32(&bars->full[s]), 1);
            mbar_init(smem_u32(&bars->empty[s]), 1);
        }
        mbar_init(smem_u32(&bars->tmem_full), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {  // one warp allocates the accumulator's TMEM columns and gives the permit back
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer: the four split operand tiles of a k-block travel once and feed three products =====
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kStages, round = kb / kStages;
                if (round > 0) mbar_wait(smem_u32(&bars->empty[s]), (round - 1) & 1);
                const uint32_t full = smem_u32(&bars->full[s]), st = sa0 + s * kStageBytes;
                mbar_expect_tx(full, kStageBytes);
                tma_load_2d(st, &map_ahi, full, kb * BK, m0);
                tma_load_2d(st + kTileBytes, &map_alo, full, kb * BK, m0);
                tma_load_2d(st + 2 * kTileBytes, &map_bhi, full, kb * BK, n0);
                tma_load_2d(st + 3 * kTileBytes, &map_blo, full, kb * BK, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer: per k-block lo*hi, hi*lo, hi*hi, each as BK / 8 instructions of K = 8 =====
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kStages, round = kb / kStages;
                mbar_wait(smem_u32(&bars->full[s]), round & 1);
                tc_fence_after();
                const uint32_t st = sa0 + s * kStageBytes;
                const uint64_t ahi = make_desc(st), alo = make_desc(st + kTileBytes), bhi = make_desc(st + 2 * kTileBytes), blo = make_desc(st + 3 * kTileBytes);
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) {  // UMMA K = 8 tf32 = 32 bytes: the descriptor's start address advances by 2 (16-byte units)
                    // the small products have their own accumulator: added into the running hi*hi sum one by one they
                    // would each be rounded at the magnitude of the large sum (measured: 2.5x the error)
                    umma_tf32(tmem + BN, alo + 2 * k, bhi + 2 * k, kIdesc, (kb | k) ? 1u : 0u);
                    umma_tf32(tmem + BN, ahi + 2 * k, blo + 2 * k, kIdesc, 1u);
                    umma_tf32(tmem, ahi + 2 * k, bhi + 2 * k, kIdesc, (kb | k) ? 1u : 0u);
                }
                umma_commit(smem_u32(&bars->empty[s]));  // the stage is free once these MMAs have read it
            }
            umma_commit(smem_u32(&bars->tmem_full));     // the accumulator is complete once every MMA has retired
        }
    } else {
        // ===== epilogue: warp w reads TMEM lanes 32 (w % 4) .. +31 = output rows m0 + 32 (w % 4) + lane =====
        const int q = warp & 3;
        // The accumulator arrives one ROW per lane (32 consecutive columns in 32 registers); C is row-major, so the tile
        // goes through shared memory once -- the operand ring is free by then -- and leaves as full 128-byte row segments:
        // 8 lanes per row, 4 rows per instruction.  The C values of a 32-column chunk are loaded BEFORE they are needed
        // (chunk 0 under the main loop, chunk c+1 under the arithmetic of chunk c): a load issued behind a store to C
        // cannot be hoisted by the compiler (possible alias), and 32 dependent DRAM round trips per warp made the first
        // version's epilogue 25 us per tile.
        constexpr int kPitch = 144;  // 32 floats + 16 bytes: conflict-free 16-byte accesses by rows and by row segments
        const uint32_t stage = sa0 + (uint32_t)q * (32 * kPitch);
        const int rsub = lane >> 3, cseg = lane & 7;
        const bool full_n = (n0 + BN <= N);
        float4 cv[8], cn[8];
        auto load_chunk = [&](int c, float4 (&dst)[8]) {
            const int col = n0 + c * 32 + 4 * cseg;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = m0 + q * 32 + rsub + 4 * i;
                dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row < M && (full_n || col + 3 < N)) dst[i] = *reinterpret_cast<const float4*>(C + (long long)row * ldc + col);
            }
        };
        load_chunk(0, cv);
        mbar_wait(smem_u32(&bars->tmem_full), 0);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t v[32], w[32];
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
            tmem_ld32(v, taddr);
            tmem_ld32(w, taddr + BN);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
#pragma unroll
            for (int j = 0; j < 8; ++j)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + (uint32_t)lane * kPitch + j * 16), "r"(v[4 * j]), "r"(v[4 * j + 1]),
                             "r"(v[4 * j + 2]), "r"(v[4 * j + 3]) : "memory");
            __syncwarp();
            if (c + 1 < BN / 32) load_chunk(c + 1, cn);  // in flight under the stores below
            const int col = n0 + c * 32 + 4 * cseg;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = rsub + 4 * i;
                const int row = m0 + q * 32 + r;
                float4 d;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(d.x), "=f"(d.y), "=f"(d.z), "=f"(d.w) : "r"(stage + (uint32_t)r * kPitch + cseg * 16) : "memory");
                if (row < M) {
                    float* cp = C + (long long)row * ldc + col;
                    if (full_n || col + 3 < N) {
                        *reinterpret_cast<float4*>(cp) = make_float4(cv[i].x - d.x, cv[i].y - d.y, cv[i].z - d.z, cv[i].w - d.w);
                    } else {
                        const float dv[4] = {d.x, d.y, d.z, d.w};
                        for (int e = 0; e < 4; ++e)
                            if (col + e < N) cp[e] -= dv[e];
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) cv[i] = cn[i];
            __syncwarp();  // the staging rows are rewritten by the next 32 columns
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

// A (m x k, row-major, lda) -> hi / lo TF32 terms, packed m x k
__global__ void split_a_kernel(const float* __restrict__ A, long long lda, float* __restrict__ hi, float* __restrict__ lo, long long m, int k) {
    const long long total = m * k;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / k;
        const int c = (int)(i - r * k);
        const float a = A[r * lda + c];
        uint32_t h, l;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(a));
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(a - __uint_as_float(h)));
        hi[i] = __uint_as_float(h);
        lo[i] = __uint_as_float(l);
    }
}

// B (k x n, row-major, ldb) -> transposed hi / lo terms, packed n x k (K-major operand for the MMA)
__global__ void __launch_bounds__(256) split_bt_kernel(const float* __restrict__ B, long long ldb, float* __restrict__ hiT, float* __restrict__ loT, int k, long long n) {
    __shared__ float tile[32][33];
    const long long n0 = (long long)blockIdx.x * 32;
    const int k0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int kk = k0 + r;
        const long long nn = n0 + tx;
        tile[r][tx] = (kk < k && nn < n) ? B[(long long)kk * ldb + nn] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const long long nn = n0 + r;
        const int kk = k0 + tx;
        if (nn < n && kk < k) {
            const float b = tile[tx][r];
            uint32_t h, l;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(b));
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(b - __uint_as_float(h)));
            hiT[nn * k + kk] = __uint_as_float(h);
            loT[nn * k + kk] = __uint_as_float(l);
        }
    }
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                              const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode(EncodeFn* out) {
    static EncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        LAIR_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        LAIR_REQUIRE(p != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available from this driver");
        fn = reinterpret_cast<EncodeFn>(p);
    }
    *out = fn;
    return LAIR_B200_OK;
}

// rows x k f32, packed (pitch k), boxes of 128 rows x 16 columns, 64B swizzle; out-of-range rows read as zero
int make_map(EncodeFn enc, CUtensorMap* map, const float* ptr, int64_t rows, int64_t k) {
    const cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)k * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    LAIR_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for a %lld x %lld operand", (int)r, (long long)rows, (long long)k);
    return LAIR_B200_OK;
}

}  // namespace

bool sgemm_tf32x3_supported(int64_t m, int64_t n, int64_t k, const float* d_c, int64_t ldc) {
    return m >= 256 && n >= 256 && k >= 32 && k % 32 == 0 && k <= 1024 && ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(d_c) % 16 == 0);
}

// C (m x n) -= A (m x k) * B (k x n), all row-major f32.  `slot` picks the workspace (one per calling stream of the sweep).
int sgemm_tf32x3_minus_dev(int64_t m, int64_t n, int64_t k, const float* d_a, int64_t lda, const float* d_b, int64_t ldb, float* d_c, int64_t ldc,
                           int slot, cudaStream_t s) {
    LAIR_REQUIRE(sgemm_tf32x3_supported(m, n, k, d_c, ldc), "sgemm_tf32x3: unsupported shape or alignment");
    EncodeFn enc;
    LAIR_CHECK(get_encode(&enc));
    void* ws = nullptr;
    const size_t a_elems = (size_t)m * k, b_elems = (size_t)n * k;
    LAIR_CHECK(ensure_work(slot, (2 * a_elems + 2 * b_elems) * sizeof(float) + 1024, &ws, s));
    float* ahi = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
    float* alo = ahi + a_elems;
    float* bhi = alo + a_elems;
    float* blo = bhi + b_elems;
    {
        const long long total = (long long)a_elems;
        int grid = (int)((total + 255) / 256);
        if (grid > 148 * 16) grid = 148 * 16;
        split_a_kernel<<<grid, 256, 0, s>>>(d_a, (long long)lda, ahi, alo, (long long)m, (int)k);
        LAIR_LAUNCH_CHECK();
        dim3 g2((unsigned)((n + 31) / 32), (unsigned)((k + 31) / 32));
        split_bt_kernel<<<g2, 256, 0, s>>>(d_b, (long long)ldb, bhi, blo, (int)k, (long long)n);
        LAIR_LAUNCH_CHECK();
    }
    CUtensorMap mah, mal, mbh, mbl;
    LAIR_CHECK(make_map(enc, &mah, ahi, m, k));
    LAIR_CHECK(make_map(enc, &mal, alo, m, k));
    LAIR_CHECK(make_map(enc, &mbh, bhi, n, k));
    LAIR_CHECK(make_map(enc, &mbl, blo, n, k));
    constexpr size_t kSmem = kStages * kStageBytes + sizeof(Bars) + 1024;
    static KernCfg kc;
    if (stale_for_context(kc.epoch)) kc.devmask = 0;
    int dev = 0;
    LAIR_CUDA_CHECK(cudaGetDevice(&dev));
    if (!((kc.devmask >> dev) & 1u)) {
        LAIR_CUDA_CHECK(cudaFuncSetAttribute(sgemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
        kc.devmask |= 1u << dev;
    }
    const int64_t tiles_m = (m + BM - 1) / BM, tiles_n = (n + BN - 1) / BN;
    LAIR_REQUIRE(tiles_m * tiles_n < (1ll << 31), "sgemm_tf32x3: too many tiles");
    ProfScope prof(kProfGemm, s, 2.0 * (double)m * (double)n * (double)k);
    sgemm_tf32x3_kernel<<<(unsigned)(tiles_m * tiles_n), kThreads, kSmem, s>>>(mah, mal, mbh, mbl, d_c, (long long)ldc, (int)m, (int)n, (int)k, (int)tiles_m);
    LAIR_LAUNCH_CHECK();
    return LAIR_B200_OK;
}

}  // namespace lair
