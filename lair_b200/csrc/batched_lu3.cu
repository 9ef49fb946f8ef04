// Batched 32x32 LU, third generation: the same algorithm and the same arithmetic as
// batched_lu.cu (one warp per matrix, one row per lane, the reference's row-major loop
// operation for operation: src/lapack/getrf.rs:46-120, src/blas/iamax.rs:6-21 -- L\U, pivots
// and info stay BIT-IDENTICAL to the reference), rebuilt around the instruction count.
//
// ncu on the first kernel (profiles/r1_batched_ncu.md): DRAM traffic == algorithmic bytes, f32
// issue slots 60 % busy with 2 878 warp instructions per matrix, of which only ~1 000 were the
// rounded multiplies / subtracts; the rest was divergence bookkeeping (BSSY/BSYNC/BRA, the
// convergence guards ptxas puts before every warp collective when the loop trip count is not
// provably warp-uniform), register moves and scalar compares.  This kernel therefore
//   * runs ONE WARP PER CTA, so the matrix loop depends only on blockIdx (provably uniform:
//     no convergence guards, uniform registers and branches for the collectives' results);
//   * keeps the pivot search and bookkeeping branch-free: dead / live / pivot rows are told apart
//     by predicates, the singular step (max == 0) is folded in with selects; ONE divergent region
//     per column remains, around the rank-1 update of the live rows (PTX-level predication was
//     tried: ptxas lowers it to unconditional math + one SEL per register, twice the work);
//   * f32: updates TWO columns per instruction with the packed f32x2 pipe.  ptxas contracts
//     mul.rn.f32x2 + sub.rn.f32x2 into FFMA2 (one rounding -- wrong), so the product is formed
//     as fma.rn.f32x2(l, u, -0.0) with the -0.0 held in a kernel parameter: adding -0.0 is exact
//     for every product (including both zeros), and ptxas cannot fold what it cannot see;
//   * records pivots with one predicated shared-memory store instead of a per-lane select.
#include "common.cuh"
#include "pivot_key.cuh"

namespace lair {
namespace {

using u64 = unsigned long long;
constexpr unsigned kAll = 0xffffffffu;

__device__ __forceinline__ void cpa16(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__device__ __forceinline__ unsigned lo32(u64 v) { return (unsigned)v; }
__device__ __forceinline__ unsigned hi32(u64 v) { return (unsigned)(v >> 32); }
__device__ __forceinline__ u64 pack32(unsigned lo, unsigned hi) { return ((u64)hi << 32) | lo; }

// a (one packed pair = 2 columns) -= l * u: the product and the difference are rounded separately.
// The product is fma(l, u, -0.0) with the -0.0 pair in a register ptxas cannot see through (see above).
__device__ __forceinline__ void sub_mul_f32x2(u64& a, u64 u, u64 ll, u64 negzero) {
    u64 t;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(t) : "l"(ll), "l"(u), "l"(negzero));
    asm("sub.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(t));
}

// ------------------------------------------------------------------------------------------
// f32: 16 packed column pairs per lane
// ------------------------------------------------------------------------------------------
template <int MINB>
__global__ void __launch_bounds__(32, MINB)
batched_lu32_v3_f32(float* __restrict__ A, int32_t* __restrict__ ipiv, int32_t* __restrict__ info, long long batch, u64 negzero) {
    constexpr int N = 32, LD = N + 4;  // 144-byte row pitch: conflict-free 128-bit row accesses
    __shared__ __align__(16) float mat[N * LD];
    __shared__ __align__(16) float rowbuf[2][LD];
    __shared__ int pivs[N];
    const int lane = threadIdx.x;

    for (long long mi = blockIdx.x; mi < batch; mi += gridDim.x) {
        float* g = A + mi * (long long)(N * N);
        if (mi + gridDim.x < batch) {  // next matrix of this CTA into L2 while this one is factored
            const char* nxt = reinterpret_cast<const char*>(A + (mi + gridDim.x) * (long long)(N * N));
            asm volatile("prefetch.global.L2 [%0];\n" ::"l"(nxt + lane * 128));
        }
#pragma unroll
        for (int c = lane; c < N * 8; c += 32) cpa16(mat + (c >> 3) * LD + (c & 7) * 4, g + (size_t)c * 4);
        cpa_wait();
        __syncwarp();
        u64 ap[16];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(mat + lane * LD + c * 4);
            ap[2 * c] = v.x;
            ap[2 * c + 1] = v.y;
        }
        __syncwarp();

        int pos = lane;  // current logical row of the row this lane owns
        int alive = 1;   // not yet a pivot row
        int sing = -1;

#pragma unroll
        for (int j = 0; j < N; ++j) {
            // -- iamax over the live rows (iamax.rs:10-19): |x| bits order like unsigned; zero and NaN -> 0 --
            const unsigned xb = (j & 1) ? hi32(ap[j >> 1]) : lo32(ap[j >> 1]);
            const unsigned mag = xb & 0x7fffffffu;
            const unsigned key = (alive && mag <= 0x7f800000u) ? mag : 0u;
            const unsigned kmax = __reduce_max_sync(kAll, key);
            const bool nonzero = kmax != 0u;  // max_val == 0: singular step, no swap, no update (getrf.rs:72-73)
            // strict `>` in the reference == lowest logical row among equal maxima
            const unsigned ppos = __reduce_min_sync(kAll, (nonzero && key == kmax) ? (unsigned)pos : 0xffffffffu);
            const bool is_w = (unsigned)pos == ppos;
            if (lane == 0) pivs[j] = nonzero ? (int)ppos : j;
            sing = nonzero ? sing : j;
            pos = (nonzero && pos == j) ? (int)ppos : pos;  // the row sitting at j moves to the pivot's old place
            pos = is_w ? j : pos;                           // the pivot row moves to j
            alive = (pos > j) ? alive : 0;  // the row now at j is done (pivot, or left in place by a singular step)
            // -- the pivot row (columns >= j) through shared memory --
            float* rb = rowbuf[j & 1];
            if (is_w) {
#pragma unroll
                for (int c = j / 4; c < 8; ++c) *reinterpret_cast<ulonglong2*>(rb + c * 4) = make_ulonglong2(ap[2 * c], ap[2 * c + 1]);
            }
            __syncwarp();
            if (alive && nonzero) {
                const float recip = __frcp_rn(rb[j]);  // A::one() / pivot (getrf.rs:76), correctly rounded == 1.0 / x
                const float l = __fmul_rn(__uint_as_float(xb), recip);  // *row_j *= pivot_recip (getrf.rs:81)
                const unsigned lb = __float_as_uint(l);
                const u64 ll = pack32(lb, lb);
                // columns j+1 .. 31 (getrf.rs:86-87): the odd column next to j alone, then packed pairs
                if ((j & 1) == 0) {
                    const float x = __fsub_rn(__uint_as_float(hi32(ap[j >> 1])), __fmul_rn(l, rb[j + 1]));
                    ap[j >> 1] = pack32(lb, __float_as_uint(x));
                } else {
                    ap[j >> 1] = pack32(lo32(ap[j >> 1]), lb);
                }
                const int p0 = (j >> 1) + 1;  // first whole pair right of column j
                if ((p0 & 1) && p0 < 16)      // an unaligned leading pair: second half of its 128-bit chunk
                    sub_mul_f32x2(ap[p0], *reinterpret_cast<const u64*>(rb + 2 * p0), ll, negzero);
#pragma unroll
                for (int c = (p0 + 1) / 2; c < 8; ++c) {
                    const ulonglong2 uu = *reinterpret_cast<const ulonglong2*>(rb + c * 4);
                    sub_mul_f32x2(ap[2 * c], uu.x, ll, negzero);
                    sub_mul_f32x2(ap[2 * c + 1], uu.y, ll, negzero);
                }
            }
        }

        // ---- rows to their final positions, then coalesced store ----
#pragma unroll
        for (int c = 0; c < 8; ++c) *reinterpret_cast<ulonglong2*>(mat + pos * LD + c * 4) = make_ulonglong2(ap[2 * c], ap[2 * c + 1]);
        __syncwarp();
#pragma unroll
        for (int c = lane; c < N * 8; c += 32)
            *reinterpret_cast<float4*>(g + (size_t)c * 4) = *reinterpret_cast<const float4*>(mat + (c >> 3) * LD + (c & 7) * 4);
        ipiv[mi * N + lane] = pivs[lane];
        if (lane == 0) info[mi] = sing;
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// f64
// ------------------------------------------------------------------------------------------
template <int MINB>
__global__ void __launch_bounds__(32, MINB)
batched_lu32_v3_f64(double* __restrict__ A, int32_t* __restrict__ ipiv, int32_t* __restrict__ info, long long batch) {
    constexpr int N = 32, LD = N + 2;  // 272-byte row pitch
    __shared__ __align__(16) double mat[N * LD];
    __shared__ __align__(16) double rowbuf[2][LD];
    __shared__ int pivs[N];
    const int lane = threadIdx.x;

    for (long long mi = blockIdx.x; mi < batch; mi += gridDim.x) {
        double* g = A + mi * (long long)(N * N);
        if (mi + gridDim.x < batch) {
            const char* nxt = reinterpret_cast<const char*>(A + (mi + gridDim.x) * (long long)(N * N));
            asm volatile("prefetch.global.L2 [%0];\n" ::"l"(nxt + lane * 128));
            asm volatile("prefetch.global.L2 [%0];\n" ::"l"(nxt + 4096 + lane * 128));
        }
#pragma unroll
        for (int c = lane; c < N * 16; c += 32) cpa16(mat + (c >> 4) * LD + (c & 15) * 2, g + (size_t)c * 2);
        cpa_wait();
        __syncwarp();
        double a[N];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const double2 v = *reinterpret_cast<const double2*>(mat + lane * LD + c * 2);
            a[2 * c] = v.x;
            a[2 * c + 1] = v.y;
        }
        __syncwarp();

        int pos = lane;
        int alive = 1;
        int sing = -1;

#pragma unroll
        for (int j = 0; j < N; ++j) {
            const u64 xb = (u64)__double_as_longlong(a[j]);
            const u64 mag = xb & 0x7fffffffffffffffull;
            const bool ok = alive && mag <= 0x7ff0000000000000ull;
            const unsigned khi = ok ? (unsigned)(mag >> 32) : 0u;
            const unsigned mh = __reduce_max_sync(kAll, khi);
            const unsigned klo = (ok && khi == mh) ? (unsigned)mag : 0u;
            const unsigned ml = __reduce_max_sync(kAll, klo);
            const bool nonzero = (mh | ml) != 0u;
            const bool cand = ok && khi == mh && (unsigned)mag == ml;
            const unsigned ppos = __reduce_min_sync(kAll, (nonzero && cand) ? (unsigned)pos : 0xffffffffu);
            const bool is_w = (unsigned)pos == ppos;
            if (lane == 0) pivs[j] = nonzero ? (int)ppos : j;
            sing = nonzero ? sing : j;
            pos = (nonzero && pos == j) ? (int)ppos : pos;
            pos = is_w ? j : pos;
            alive = (pos > j) ? alive : 0;  // the row now at j is done (pivot, or left in place by a singular step)
            double* rb = rowbuf[j & 1];
            if (is_w) {
#pragma unroll
                for (int c = j / 2; c < 16; ++c) *reinterpret_cast<double2*>(rb + c * 2) = make_double2(a[2 * c], a[2 * c + 1]);
            }
            __syncwarp();
            if (alive && nonzero) {
                const double recip = __drcp_rn(rb[j]);  // correctly rounded == 1.0 / x (getrf.rs:76)
                const double l = __dmul_rn(a[j], recip);
                a[j] = l;
                if ((j & 1) == 0) a[j + 1] = __dsub_rn(a[j + 1], __dmul_rn(l, rb[j + 1]));
#pragma unroll
                for (int c = j / 2 + 1; c < 16; ++c) {
                    const double2 uu = *reinterpret_cast<const double2*>(rb + c * 2);
                    a[2 * c] = __dsub_rn(a[2 * c], __dmul_rn(l, uu.x));
                    a[2 * c + 1] = __dsub_rn(a[2 * c + 1], __dmul_rn(l, uu.y));
                }
            }
        }

#pragma unroll
        for (int c = 0; c < 16; ++c) *reinterpret_cast<double2*>(mat + pos * LD + c * 2) = make_double2(a[2 * c], a[2 * c + 1]);
        __syncwarp();
#pragma unroll
        for (int c = lane; c < N * 16; c += 32)
            *reinterpret_cast<double2*>(g + (size_t)c * 2) = *reinterpret_cast<const double2*>(mat + (c >> 4) * LD + (c & 15) * 2);
        ipiv[mi * N + lane] = pivs[lane];
        if (lane == 0) info[mi] = sing;
        __syncwarp();
    }
}

template <class K>
int launch_v3(K kern, int& blocks_per_sm, bool& configured) {
    if (!configured) {
        LAIR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, 32, 0));
        if (blocks_per_sm < 1) blocks_per_sm = 1;
        configured = true;
    }
    return LAIR_B200_OK;
}

}  // namespace

// Full 32 x 32, 16-byte aligned batches only (the caller checks).  variant: CTAs per SM hint (0 = default).
template <>
int getrf_batched32v3_dev<float>(int64_t batch, float* d_a, int32_t* d_ipiv, int32_t* d_info, int variant, cudaStream_t s) {
    auto kern = variant == 1 ? batched_lu32_v3_f32<16> : batched_lu32_v3_f32<24>;
    static int bps[2] = {0, 0};
    static bool conf[2] = {false, false};
    const int v = variant == 1 ? 1 : 0;
    LAIR_CHECK(launch_v3(kern, bps[v], conf[v]));
    const long long cap = (long long)ctx().sm_count * bps[v];
    const int grid = (int)(batch < cap ? batch : cap);
    if (grid < 1) return LAIR_B200_OK;
    ProfScope prof(kProfBatched, s, (double)batch * (2.0 * 32 * 32 * sizeof(float) + 4.0 * 32));
    const u64 negzero = 0x8000000080000000ull;
    kern<<<grid, 32, 0, s>>>(d_a, d_ipiv, d_info, (long long)batch, negzero);
    LAIR_LAUNCH_CHECK();
    return LAIR_B200_OK;
}

template <>
int getrf_batched32v3_dev<double>(int64_t batch, double* d_a, int32_t* d_ipiv, int32_t* d_info, int variant, cudaStream_t s) {
    auto kern = variant == 1 ? batched_lu32_v3_f64<12> : batched_lu32_v3_f64<16>;
    static int bps[2] = {0, 0};
    static bool conf[2] = {false, false};
    const int v = variant == 1 ? 1 : 0;
    LAIR_CHECK(launch_v3(kern, bps[v], conf[v]));
    const long long cap = (long long)ctx().sm_count * bps[v];
    const int grid = (int)(batch < cap ? batch : cap);
    if (grid < 1) return LAIR_B200_OK;
    ProfScope prof(kProfBatched, s, (double)batch * (2.0 * 32 * 32 * sizeof(double) + 4.0 * 32));
    kern<<<grid, 32, 0, s>>>(d_a, d_ipiv, d_info, (long long)batch);
    LAIR_LAUNCH_CHECK();
    return LAIR_B200_OK;
}

}  // namespace lair
