// Batched 32x32 LU, second layout: TWO matrices per warp (one per half-warp), two matrix rows
// per lane (rows l and l+16).  Same arithmetic as batched_lu.cu -- the reference's row-major
// loop operation for operation (src/lapack/getrf.rs:46-120, src/blas/iamax.rs:6-21), so L\U,
// pivots and info stay BIT-IDENTICAL to the reference -- but every non-arithmetic warp
// instruction of a step (arg-max reductions, position bookkeeping, the pivot-row broadcast
// through shared memory, the reciprocal) now serves two matrices, and the two row slots give
// the scheduler two independent update chains per lane.  ncu on the one-matrix-per-warp kernel
// (profiles/r1_batched_ncu.md) showed it issue/latency-bound with the shared-memory pipe
// throttling, not DRAM-bound; this layout halves the per-matrix instruction overhead.
#include "common.cuh"
#include "pivot_key.cuh"

namespace lair {
namespace {

__device__ __forceinline__ void cp_async16z(void* smem_dst, const void* gmem_src, int bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all2() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

template <class T> struct Vec16b;
template <> struct Vec16b<float> { using type = float4; static constexpr int n = 4; };
template <> struct Vec16b<double> { using type = double2; static constexpr int n = 2; };

// max of a pivot key over the 16 lanes selected by `mask`
template <class KT>
__device__ __forceinline__ KT half_max(unsigned mask, KT k) {
    if (sizeof(KT) == 8) {
        const uint32_t hi = (uint32_t)((unsigned long long)k >> 32);
        const uint32_t mh = __reduce_max_sync(mask, hi);
        const uint32_t lo = (hi == mh) ? (uint32_t)k : 0u;
        const uint32_t ml = __reduce_max_sync(mask, lo);
        return (KT)(((unsigned long long)mh << 32) | ml);
    } else {
        return (KT)__reduce_max_sync(mask, (uint32_t)k);
    }
}

template <class T, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
batched_lu32x2_kernel(T* __restrict__ A, int32_t* __restrict__ ipiv, int32_t* __restrict__ info, long long batch) {
    constexpr int N = 32;
    using V = typename Vec16b<T>::type;
    constexpr int VEC = Vec16b<T>::n;
    constexpr int LD = N + VEC;   // padded row: conflict-free 128-bit row accesses
    constexpr int CPR = N / VEC;  // 16-byte chunks per row
    using K = PivotKey<T>;
    using KT = typename K::type;
    using O = Ops<T>;
    constexpr unsigned NOPOS = 0x7fffffffu;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = lane >> 4, sl = lane & 15;
    const unsigned hmask = 0xffffu << (16 * h);
    T* mats = reinterpret_cast<T*>(smem_raw) + (size_t)warp * (2 * N * LD + 4 * LD);
    T* rowbuf = mats + 2 * N * LD;  // [half][parity][LD]

    const long long npairs = (batch + 1) / 2;
    const long long warp_global = (long long)blockIdx.x * WARPS + warp;
    const long long warp_total = (long long)gridDim.x * WARPS;

    for (long long pi = warp_global; pi < npairs; pi += warp_total) {
        const long long m = 2 * pi + h;  // this half-warp's matrix
        const bool valid = m < batch;
        T* g = A + 2 * pi * (long long)(N * N);  // the pair is contiguous in memory
        if (pi + warp_total < npairs) {  // pull the next pair into L2 while this one is factored
            const char* nxt = reinterpret_cast<const char*>(A + 2 * (pi + warp_total) * (long long)(N * N));
            for (int off = lane * 128; off < (int)(2 * N * N * sizeof(T)); off += 32 * 128)
                asm volatile("prefetch.global.L2 [%0];\n" ::"l"(nxt + off));
        }
        // ---- stage both matrices: coalesced global -> padded shared ----
#pragma unroll
        for (int c = lane; c < 2 * N * CPR; c += 32) {
            const int which = c / (N * CPR), r = (c % (N * CPR)) / CPR, cc = c % CPR;
            const bool ok = 2 * pi + which < batch;
            cp_async16z(mats + (which * N + r) * LD + cc * VEC, ok ? g + (size_t)c * VEC : g, ok ? 16 : 0);
        }
        cp_async_wait_all2();
        __syncwarp();
        T a0[N], a1[N];
#pragma unroll
        for (int c = 0; c < CPR; ++c) {
            const V v0 = *reinterpret_cast<const V*>(mats + (h * N + sl) * LD + c * VEC);
            const V v1 = *reinterpret_cast<const V*>(mats + (h * N + sl + 16) * LD + c * VEC);
            const T* p0 = reinterpret_cast<const T*>(&v0);
            const T* p1 = reinterpret_cast<const T*>(&v1);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                a0[c * VEC + e] = p0[e];
                a1[c * VEC + e] = p1[e];
            }
        }
        if (!valid) {  // odd batch: the idle half factors an identity (never stored)
#pragma unroll
            for (int k = 0; k < N; ++k) {
                a0[k] = (k == sl) ? O::one() : O::zero();
                a1[k] = (k == sl + 16) ? O::one() : O::zero();
            }
        }
        __syncwarp();

        int pos0 = sl, pos1 = sl + 16;  // logical rows of the two rows this lane owns
        int piv0 = sl, piv1 = sl + 16;  // lane sl records ipiv[sl] and ipiv[sl+16]
        int sing = -1;

#pragma unroll
        for (int j = 0; j < N; ++j) {
            // -- iamax over logical rows >= j of this half-warp's matrix (iamax.rs:10-19) --
            const bool live0 = pos0 >= j, live1 = pos1 >= j;
            const KT k0 = live0 ? K::of(a0[j]) : (KT)0;
            const KT k1 = live1 ? K::of(a1[j]) : (KT)0;
            const unsigned p0 = live0 ? (unsigned)pos0 : NOPOS, p1 = live1 ? (unsigned)pos1 : NOPOS;
            const bool take1 = k1 > k0 || (k1 == k0 && p1 < p0);
            const KT bk = take1 ? k1 : k0;
            const unsigned bp = take1 ? p1 : p0;
            const KT kmax = half_max<KT>(hmask, bk);
            // strict `>` in the reference == lowest logical row among equal maxima
            const unsigned ppos = __reduce_min_sync(hmask, (bk == kmax) ? bp : 0xffffffffu);
            const bool act = (kmax != 0);  // max_val == 0: singular step, no swap, no update (getrf.rs:72-73)
            if (!act) sing = j;
            const bool w0 = act && live0 && ((unsigned)pos0 == ppos);
            const bool w1 = act && live1 && ((unsigned)pos1 == ppos);
            if (act) {
                if (j < 16) {
                    if (sl == j) piv0 = (int)ppos;
                } else {
                    if (sl == j - 16) piv1 = (int)ppos;
                }
                const bool was_j0 = (pos0 == j), was_j1 = (pos1 == j);
                if (was_j0) pos0 = (int)ppos;  // the row sitting at j moves to the pivot's old place
                if (was_j1) pos1 = (int)ppos;
                if (w0) pos0 = j;              // the pivot row moves to j
                if (w1) pos1 = j;
            }
            // -- broadcast the pivot row (columns >= j) through shared memory --
            const int c0 = j / VEC;
            T* rb = rowbuf + (h * 2 + (j & 1)) * LD;
            if (w0) {
#pragma unroll
                for (int c = c0; c < CPR; ++c) {
                    V v;
                    T* pv = reinterpret_cast<T*>(&v);
#pragma unroll
                    for (int e = 0; e < VEC; ++e) pv[e] = a0[c * VEC + e];
                    *reinterpret_cast<V*>(rb + c * VEC) = v;
                }
            }
            if (w1) {
#pragma unroll
                for (int c = c0; c < CPR; ++c) {
                    V v;
                    T* pv = reinterpret_cast<T*>(&v);
#pragma unroll
                    for (int e = 0; e < VEC; ++e) pv[e] = a1[c * VEC + e];
                    *reinterpret_cast<V*>(rb + c * VEC) = v;
                }
            }
            __syncwarp();
            const T recip = O::recip(rb[j]);  // A::one() / pivot (getrf.rs:76); unused when !act
            const bool u0 = act && pos0 > j, u1 = act && pos1 > j;
            const T l0 = O::mul(a0[j], recip), l1 = O::mul(a1[j], recip);  // *row_j *= pivot_recip (getrf.rs:81)
            if (u0) a0[j] = l0;
            if (u1) a1[j] = l1;
#pragma unroll
            for (int c = (j + 1) / VEC; c < CPR; ++c) {
                const V v = *reinterpret_cast<const V*>(rb + c * VEC);
                const T* pv = reinterpret_cast<const T*>(&v);
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    const int k = c * VEC + e;
                    if (k > j) {
                        const T t0 = O::sub(a0[k], O::mul(l0, pv[e]));  // getrf.rs:86-87
                        const T t1 = O::sub(a1[k], O::mul(l1, pv[e]));
                        if (u0) a0[k] = t0;
                        if (u1) a1[k] = t1;
                    }
                }
            }
        }

        // ---- rows to their final positions, then coalesced store ----
#pragma unroll
        for (int c = 0; c < CPR; ++c) {
            V v0, v1;
            T* q0 = reinterpret_cast<T*>(&v0);
            T* q1 = reinterpret_cast<T*>(&v1);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                q0[e] = a0[c * VEC + e];
                q1[e] = a1[c * VEC + e];
            }
            *reinterpret_cast<V*>(mats + (h * N + pos0) * LD + c * VEC) = v0;
            *reinterpret_cast<V*>(mats + (h * N + pos1) * LD + c * VEC) = v1;
        }
        __syncwarp();
#pragma unroll
        for (int c = lane; c < 2 * N * CPR; c += 32) {
            const int which = c / (N * CPR), r = (c % (N * CPR)) / CPR, cc = c % CPR;
            if (2 * pi + which < batch)
                *reinterpret_cast<V*>(g + (size_t)c * VEC) = *reinterpret_cast<const V*>(mats + (which * N + r) * LD + cc * VEC);
        }
        if (valid) {
            ipiv[m * N + sl] = piv0;
            ipiv[m * N + 16 + sl] = piv1;
            if (sl == 0) info[m] = sing;
        }
        __syncwarp();
    }
}

template <class T, int WARPS, int MINB>
int launch_batched2(long long batch, T* d_a, int32_t* d_ipiv, int32_t* d_info, cudaStream_t s) {
    constexpr int N = 32, VEC = Vec16b<T>::n, LD = N + VEC;
    auto kern = batched_lu32x2_kernel<T, WARPS, MINB>;
    const size_t smem = (size_t)WARPS * (2 * N * LD + 4 * LD) * sizeof(T);
    static bool configured = false;
    static int blocks_per_sm = 1;
    if (!configured) {
        LAIR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        LAIR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, WARPS * 32, smem));
        if (blocks_per_sm < 1) blocks_per_sm = 1;
        configured = true;
    }
    const long long npairs = (batch + 1) / 2;
    const long long want = (npairs + WARPS - 1) / WARPS;
    const long long cap = (long long)ctx().sm_count * blocks_per_sm;
    const int grid = (int)(want < cap ? want : cap);
    if (grid < 1) return LAIR_B200_OK;
    ProfScope prof(kProfBatched, s, (double)batch * (2.0 * N * N * sizeof(T) + 4.0 * N));
    kern<<<grid, WARPS * 32, smem, s>>>(d_a, d_ipiv, d_info, batch);
    LAIR_LAUNCH_CHECK();
    return LAIR_B200_OK;
}

}  // namespace

// variant: 0 = register bound for 2 CTAs (f64) / 4 CTAs (f32) of 4 warps per SM, 1 = one step tighter
template <class T>
int getrf_batched32x2_dev(int64_t batch, T* d_a, int32_t* d_ipiv, int32_t* d_info, int variant, cudaStream_t s) {
    if (sizeof(T) == 8) {
        if (variant == 1) return launch_batched2<T, 4, 3>(batch, d_a, d_ipiv, d_info, s);
        return launch_batched2<T, 4, 2>(batch, d_a, d_ipiv, d_info, s);
    }
    if (variant == 1) return launch_batched2<T, 4, 4>(batch, d_a, d_ipiv, d_info, s);
    return launch_batched2<T, 4, 3>(batch, d_a, d_ipiv, d_info, s);
}

template int getrf_batched32x2_dev<float>(int64_t, float*, int32_t*, int32_t*, int, cudaStream_t);
template int getrf_batched32x2_dev<double>(int64_t, double*, int32_t*, int32_t*, int, cudaStream_t);

}  // namespace lair
