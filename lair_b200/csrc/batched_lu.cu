// Batched LU of independent n x n (n <= 32) row-major matrices: one warp per matrix,
// one matrix row per lane, the whole matrix on chip.  HBM-bound path:
// algorithmic traffic = 2*n*n*sizeof(T) + 4*n bytes per matrix.
//
// Arithmetic follows the reference's row-major body step for step
// (src/lapack/getrf.rs:46-120): first-max pivot (src/blas/iamax.rs:6-21), reciprocal
// multiply for the multipliers, then a rounded multiply and a rounded subtract per
// element -- so L\U, ipiv and info are BIT-IDENTICAL to the reference.
//
// Row interchanges are logical: a lane keeps its row in registers for the whole
// factorization and only its position `pos` changes; rows are written to their final
// positions through shared memory at the end (coalesced 128-bit global stores).
#include "common.cuh"
#include "pivot_key.cuh"

namespace lair {
namespace {

constexpr unsigned kFull = kFullMask;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

template <class T> struct Vec16;  // 16-byte vector of T
template <> struct Vec16<float> { using type = float4; static constexpr int n = 4; };
template <> struct Vec16<double> { using type = double2; static constexpr int n = 2; };

// VAR bit 0: coarse-key arg-max (1 REDUX + vote) instead of full-key REDUX chain;
// VAR bit 1: consume the pivot row chunk by chunk from shared memory instead of a register copy.
template <class T, int WARPS, int MINB, bool FULL, int VAR>
__global__ void __launch_bounds__(WARPS * 32, MINB)
batched_lu32_kernel(T* __restrict__ A, int32_t* __restrict__ ipiv, int32_t* __restrict__ info, long long batch, int n) {
    constexpr int N = 32;
    using V = typename Vec16<T>::type;
    constexpr int VEC = Vec16<T>::n;
    constexpr int LD = N + VEC;           // padded smem row: conflict-free 128-bit row reads
    constexpr int CPR = N / VEC;          // 16-byte chunks per row
    using K = PivotKey<T>;
    using O = Ops<T>;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T* mat = reinterpret_cast<T*>(smem_raw) + (size_t)warp * (N * LD + 2 * LD);
    T* rowbuf = mat + N * LD;

    const long long warp_global = (long long)blockIdx.x * WARPS + warp;
    const long long warp_total = (long long)gridDim.x * WARPS;

    for (long long mi = warp_global; mi < batch; mi += warp_total) {
        T* g = A + mi * (long long)n * n;
        // pull this warp's NEXT matrix into L2 while the current one is factored (there is no room
        // for a second shared-memory buffer without losing resident warps)
        if (FULL && mi + warp_total < batch) {
            const char* nxt = reinterpret_cast<const char*>(A + (mi + warp_total) * (long long)N * N);
            for (int off = lane * 128; off < (int)(N * N * sizeof(T)); off += 32 * 128)
                asm volatile("prefetch.global.L2 [%0];\n" ::"l"(nxt + off));
        }
        // ---- stage the matrix: coalesced global -> padded shared ----
        if (FULL) {
#pragma unroll
            for (int c = lane; c < N * CPR; c += 32) {
                int r = c / CPR, cc = c % CPR;
                cp_async16(mat + r * LD + cc * VEC, g + (size_t)c * VEC);
            }
            cp_async_wait_all();
        } else {
            // n < 32: embed in diag(A, I) so the extra steps are no-ops
            for (int idx = lane; idx < N * N; idx += 32) {
                int r = idx >> 5, c = idx & 31;
                T v = (r == c) ? O::one() : O::zero();
                if (r < n && c < n) v = g[r * n + c];
                mat[r * LD + c] = v;
            }
        }
        __syncwarp();
        T a[N];
#pragma unroll
        for (int c = 0; c < CPR; ++c) {
            V v = *reinterpret_cast<const V*>(mat + lane * LD + c * VEC);
            const T* pv = reinterpret_cast<const T*>(&v);
#pragma unroll
            for (int e = 0; e < VEC; ++e) a[c * VEC + e] = pv[e];
        }
        __syncwarp();

        int pos = lane;      // current logical row of the row this lane owns
        int mypiv = lane;    // lane j records ipiv[j]; identity unless a swap happens (getrf.rs:18-19,62-64)
        int sing = -1;

#pragma unroll
        for (int j = 0; j < N; ++j) {
            if (!FULL && j >= n) break;
            // -- iamax over logical rows >= j (iamax.rs:10-19) --
            // strict `>` in the reference == lowest logical row among equal maxima; one REDUX + one
            // vote on the top 32 bits of |x| decide it unless lanes tie there (pivot_key.cuh)
            const bool live = pos >= j;
            const typename K::type key = live ? K::of(a[j]) : (typename K::type)0;
            typename K::type kmax;
            unsigned ppos;
            bool is_w;
            if constexpr ((VAR & 1) != 0) {
                int wl;
                warp_argmax<typename K::type>(key, live ? (unsigned)pos : 0x7fffffffu, kmax, ppos, wl);
                is_w = (lane == wl);
            } else {
                kmax = K::warp_max(key);
                const bool cand = live && (key == kmax);
                ppos = __reduce_min_sync(kFull, cand ? (unsigned)pos : 0xffffffffu);
                is_w = cand && ((unsigned)pos == ppos);
            }
            if (kmax == 0) {  // max_val == 0: singular step, no swap, no update (getrf.rs:72-73)
                sing = j;
                continue;
            }
            if (lane == j) mypiv = (int)ppos;
            if (pos == j) pos = (int)ppos;  // the row sitting at j moves to the pivot's old place
            if (is_w) pos = j;              // the pivot row moves to j
            // -- broadcast the pivot row (columns >= j) through shared memory --
            const int c0 = j / VEC;
            T* rb = rowbuf + (j & 1) * LD;
            if (is_w) {
#pragma unroll
                for (int c = c0; c < CPR; ++c) {
                    V v;
                    T* pv = reinterpret_cast<T*>(&v);
#pragma unroll
                    for (int e = 0; e < VEC; ++e) pv[e] = a[c * VEC + e];
                    *reinterpret_cast<V*>(rb + c * VEC) = v;
                }
            }
            __syncwarp();
            if constexpr ((VAR & 2) != 0) {
                const T recip = O::recip(rb[j]);  // A::one() / pivot (getrf.rs:76)
                if (pos > j) {
                    const T l = O::mul(a[j], recip);  // *row_j *= pivot_recip (getrf.rs:81)
                    a[j] = l;
                    // the pivot row is consumed chunk by chunk straight from shared memory (no register copy)
#pragma unroll
                    for (int c = (j + 1) / VEC; c < CPR; ++c) {
                        const V v = *reinterpret_cast<const V*>(rb + c * VEC);
                        const T* pv = reinterpret_cast<const T*>(&v);
#pragma unroll
                        for (int e = 0; e < VEC; ++e) {
                            const int k = c * VEC + e;
                            if (k > j) a[k] = O::sub(a[k], O::mul(l, pv[e]));  // getrf.rs:86-87
                        }
                    }
                }
            } else {
                T u[N];
#pragma unroll
                for (int c = c0; c < CPR; ++c) {
                    V v = *reinterpret_cast<const V*>(rb + c * VEC);
                    const T* pv = reinterpret_cast<const T*>(&v);
#pragma unroll
                    for (int e = 0; e < VEC; ++e) u[c * VEC + e] = pv[e];
                }
                const T recip = O::recip(u[j]);  // A::one() / pivot (getrf.rs:76)
                if (pos > j) {
                    const T l = O::mul(a[j], recip);  // *row_j *= pivot_recip (getrf.rs:81)
                    a[j] = l;
#pragma unroll
                    for (int k = j + 1; k < N; ++k) a[k] = O::sub(a[k], O::mul(l, u[k]));  // getrf.rs:86-87
                }
            }
        }

        // ---- rows to their final positions, then coalesced store ----
#pragma unroll
        for (int c = 0; c < CPR; ++c) {
            V v;
            T* pv = reinterpret_cast<T*>(&v);
#pragma unroll
            for (int e = 0; e < VEC; ++e) pv[e] = a[c * VEC + e];
            *reinterpret_cast<V*>(mat + pos * LD + c * VEC) = v;
        }
        __syncwarp();
        if (FULL) {
#pragma unroll
            for (int c = lane; c < N * CPR; c += 32) {
                int r = c / CPR, cc = c % CPR;
                *reinterpret_cast<V*>(g + (size_t)c * VEC) = *reinterpret_cast<const V*>(mat + r * LD + cc * VEC);
            }
            ipiv[mi * N + lane] = mypiv;
        } else {
            for (int idx = lane; idx < n * n; idx += 32) {
                int r = idx / n, c = idx % n;
                g[idx] = mat[r * LD + c];
            }
            if (lane < n) ipiv[mi * n + lane] = mypiv;
        }
        if (lane == 0) info[mi] = sing;
        __syncwarp();
    }
}

template <class T, int WARPS, int MINB, bool FULL, int VAR>
int launch_batched(long long batch, int n, T* d_a, int32_t* d_ipiv, int32_t* d_info, cudaStream_t s) {
    constexpr int N = 32, VEC = Vec16<T>::n, LD = N + VEC;
    auto kern = batched_lu32_kernel<T, WARPS, MINB, FULL, VAR>;
    size_t smem = (size_t)WARPS * (N * LD + 2 * LD) * sizeof(T);
    static KernCfg kc;
    if (stale_for_context(kc.epoch)) kc.bps = 0, kc.devmask = 0;
    int dev = 0;
    LAIR_CUDA_CHECK(cudaGetDevice(&dev));
    if (!((kc.devmask >> dev) & 1u)) {
        LAIR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (kc.bps == 0) {
            LAIR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&kc.bps, kern, WARPS * 32, smem));
            if (kc.bps < 1) kc.bps = 1;
        }
        kc.devmask |= 1u << dev;
    }
    const int blocks_per_sm = kc.bps;
    long long want = (batch + WARPS - 1) / WARPS;
    long long cap = (long long)ctx().sm_count * blocks_per_sm;
    int grid = (int)(want < cap ? want : cap);
    if (grid < 1) return LAIR_B200_OK;
    ProfScope prof(kProfBatched, s, (double)batch * (2.0 * n * n * sizeof(T) + 4.0 * n));
    kern<<<grid, WARPS * 32, smem, s>>>(d_a, d_ipiv, d_info, batch, n);
    LAIR_LAUNCH_CHECK();
    return LAIR_B200_OK;
}

}  // namespace

template <class T>
int getrf_batched_dev(int64_t batch, int64_t n, T* d_a, int32_t* d_ipiv, int32_t* d_info, cudaStream_t s) {
    LAIR_REQUIRE(batch >= 0 && n >= 0, "batched getrf: negative size");
    LAIR_REQUIRE(n <= 32, "batched getrf supports n <= 32 (got %lld)", (long long)n);
    if (batch == 0 || n == 0) return LAIR_B200_OK;
    LAIR_REQUIRE(d_a && d_ipiv && d_info, "batched getrf: null pointer");
    const bool full = (n == 32) && (reinterpret_cast<uintptr_t>(d_a) % 16 == 0);
    // Tuning variants (option "batched_cfg", bit field): bit 0 = tighter register bound (more resident
    // warps), bit 1 = coarse-key arg-max, bit 2 = pivot row consumed from shared memory; 32 / 128 / 256
    // select the later kernels (batched_lu4.cu, batched_lu5.cu).  The default was picked on a B200.  (Round 2 removed the
    // generations that lost every A/B of round 1 -- two matrices per warp x2 (cfg 8, 64), branch-free v3 (16), look-ahead
    // pivoting v7 (130/131): numbers in profiles/r1b_batched_v4.md.)
    constexpr int kLo = sizeof(T) == 8 ? 3 : 5;
    constexpr int kHi = sizeof(T) == 8 ? 4 : 8;
    int64_t cfg = ctx().opt.batched_cfg;
    // measured best on B200 (profiles/r1b_batched_v4.md): the straight-line sixth-generation kernel --
    // cfg 129 (tighter register bound): f32 320.5 M/s (cfg 128: 310, cfg 33: 266, cfg 0: 239); f64 162.1 M/s (128: 157.4, cfg 0: 123)
    if (cfg < 0) cfg = 129;
    if (!full) return launch_batched<T, 4, kLo, false, 0>(batch, (int)n, d_a, d_ipiv, d_info, s);
    if (cfg & 512) return getrf_batched32v9_dev<T>(batch, d_a, d_ipiv, d_info, (int)(cfg & 3), s);  // two matrices per warp, merged winner stores
    if (cfg & 256) return getrf_batched32v8_dev<T>(batch, d_a, d_ipiv, d_info, (int)(cfg & 15), s);  // shuffle broadcast, no per-step shared-memory traffic
    if (cfg & 128) return getrf_batched32v6_dev<T>(batch, d_a, d_ipiv, d_info, (int)(cfg & 7), s);  // bit 1: look-ahead pivoting (f32), bit 2: 8-byte winner stores (f32)  // straight-line column loop + exact fallback
    if (cfg & 32) return getrf_batched32v4_dev<T>(batch, d_a, d_ipiv, d_info, (int)(cfg & 1), s);  // retiring rows, NaN-poisoned lanes
    switch (cfg & 7) {
        case 1: return launch_batched<T, 4, kHi, true, 0>(batch, 32, d_a, d_ipiv, d_info, s);
        case 2: return launch_batched<T, 4, kLo, true, 1>(batch, 32, d_a, d_ipiv, d_info, s);
        case 3: return launch_batched<T, 4, kHi, true, 1>(batch, 32, d_a, d_ipiv, d_info, s);
        case 4: return launch_batched<T, 4, kLo, true, 2>(batch, 32, d_a, d_ipiv, d_info, s);
        case 5: return launch_batched<T, 4, kHi, true, 2>(batch, 32, d_a, d_ipiv, d_info, s);
        case 6: return launch_batched<T, 4, kLo, true, 3>(batch, 32, d_a, d_ipiv, d_info, s);
        case 7: return launch_batched<T, 4, kHi, true, 3>(batch, 32, d_a, d_ipiv, d_info, s);
        default: return launch_batched<T, 4, kLo, true, 0>(batch, 32, d_a, d_ipiv, d_info, s);
    }
}

template int getrf_batched_dev<float>(int64_t, int64_t, float*, int32_t*, int32_t*, cudaStream_t);
template int getrf_batched_dev<double>(int64_t, int64_t, double*, int32_t*, int32_t*, cudaStream_t);

}  // namespace lair
