// Panel factorization for panels of up to 16 x 512 = 8192 rows: ONE thread-block cluster,
// candidate exchange over distributed shared memory (DSMEM) instead of global memory.
//
// Same on-chip, logical-interchange scheme as panel.cu (one matrix row of W columns per
// thread, rows never move, only their position `pos` changes), but the per-column grid-wide
// exchange -- the latency that bounds the whole factorization -- never leaves the SMs:
//   thread candidate -> warp arg-max (REDUX) -> CTA candidate ->
//   every CTA PUSHES its candidate (|pivot| key, position, reciprocal, full row) into the
//   shared memory of all CTAs of the cluster (st.shared::cluster via map_shared_rank) ->
//   one cluster barrier (arrive.release / wait.acquire) ->
//   every warp picks the same winner from local shared memory and updates from registers.
// One __syncthreads and one cluster barrier per column; no global-memory round trips.
//
// Code size matters as much as the exchange: a fully unrolled 32-column body is ~100 KB of
// SASS that is executed exactly once per launch, i.e. an instruction-cache miss stream
// (first-contact measurement: 2.5 us per column no matter how the exchange was done).  So
// the live part of the row is kept in a ROTATING register window: columns are processed in
// groups of GS, the body for one group is compiled once (static register indices 0..GS-1 for
// the group's columns, window index i for column jb+i), and after each group the window
// shifts down by GS registers.  Finished entries (multipliers of L, the U part of a row that
// became a pivot) are parked in shared memory and written to global memory at the end.
//
// Pivot rule = blas::iamax (src/blas/iamax.rs:6-21): first maximum of |x| by logical row,
// NaN never wins; multipliers by reciprocal-multiply (src/lapack/getrf.rs:76-81).
#include <climits>
#include <cooperative_groups.h>

#include "common.cuh"
#include "pivot_key.cuh"

namespace cg = cooperative_groups;

namespace lair {
namespace {

constexpr int CL_TPB = 512;
constexpr int CL_MAXC = 16;

// Debug aid: cycle stamps of the per-column phases (rank 0, thread 0), summed over columns.
// [0] candidate+park, [1] __syncthreads, [2] CTA candidate + push, [3] cluster barrier,
// [4] winner select, [5] update, [6] columns timed.  Enabled by option "panel_timing".
__device__ long long g_cl_timing[8];
#define CL_STAMP(slot)                                        \
    do {                                                      \
        if (timing && rank == 0 && tid == 0) {                \
            const long long now_ = clock64();                 \
            g_cl_timing[slot] += now_ - tprev;                \
            tprev = now_;                                     \
        }                                                     \
    } while (0)

template <class T, int W>
struct ClusterSmem {
    static constexpr int NW = CL_TPB / 32;
    static constexpr int OUT_LD = W + 1;  // odd pitch: column-j accesses of all rows are conflict free
    // layout of the dynamic shared memory block
    static constexpr size_t out_bytes = (size_t)CL_TPB * OUT_LD * sizeof(T);
    static constexpr size_t rows_bytes = (size_t)2 * CL_MAXC * W * sizeof(T);
    static constexpr size_t wrow_bytes = (size_t)NW * W * sizeof(T);
    static constexpr size_t cand_bytes = (size_t)2 * CL_MAXC * 4 * sizeof(unsigned long long);
    static constexpr size_t total = out_bytes + rows_bytes + wrow_bytes + cand_bytes + 16;
};

template <class T, int W, int GS>
__global__ void __launch_bounds__(CL_TPB, 1)
panel_cluster_kernel(T* __restrict__ A, long long lda, int M, int w, int32_t* __restrict__ ipiv, int row_base,
                     int32_t* __restrict__ info, int step_base, int timing) {
    using K = PivotKey<T>;
    using KT = typename K::type;
    using SM = ClusterSmem<T, W>;
    long long tprev = 0;
    constexpr int NW = SM::NW;
    constexpr int VEC = 16 / sizeof(T);
    constexpr int OUT_LD = SM::OUT_LD;
    struct alignas(16) V16 { T v[VEC]; };
    static_assert(CL_MAXC <= NW, "one pushing warp per peer");
    static_assert(W % GS == 0 && W % VEC == 0, "shape");

    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();

    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s_rows = reinterpret_cast<T*>(smem_raw);                                        // [2][CL_MAXC][W]
    T* s_wrow = reinterpret_cast<T*>(smem_raw + SM::rows_bytes);                       // [NW][W]
    unsigned long long* s_cand = reinterpret_cast<unsigned long long*>(smem_raw + SM::rows_bytes + SM::wrow_bytes);  // [2][CL_MAXC][4]
    T* s_out = reinterpret_cast<T*>(smem_raw + SM::rows_bytes + SM::wrow_bytes + SM::cand_bytes);  // [CL_TPB][OUT_LD]
    __shared__ KT s_wkey[NW];
    __shared__ int s_wpos[NW];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- load one row per thread into the register window (window index == column) ----
    T a[W];
    const int row = rank * CL_TPB + tid;
    int pos = row < M ? row : -1;
    const bool vec_ok = (w == W) && ((lda % VEC) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    if (row < M) {
        const T* p = A + (long long)row * lda;
        if (vec_ok) {
#pragma unroll
            for (int c = 0; c < W / VEC; ++c) {
                V16 v = *reinterpret_cast<const V16*>(p + c * VEC);
#pragma unroll
                for (int e = 0; e < VEC; ++e) a[c * VEC + e] = v.v[e];
            }
        } else {
#pragma unroll
            for (int c = 0; c < W; ++c) a[c] = (c < w) ? p[c] : T(0);
        }
    } else {
#pragma unroll
        for (int c = 0; c < W; ++c) a[c] = T(0);
    }
    T* my_out = s_out + tid * OUT_LD;
    cluster.sync();  // every CTA of the cluster is running before the first remote store

    for (int jb = 0; jb < w; jb += GS) {
        const int rem = W - jb;  // live window entries: a[0 .. rem)
#pragma unroll
        for (int c = 0; c < GS; ++c) {
            const int j = jb + c;
            if (j >= w) break;  // uniform
            const int parity = j & 1;
            if (timing && rank == 0 && tid == 0) tprev = clock64();
            T* rows_p = s_rows + parity * (CL_MAXC * W);
            unsigned long long* cand_p = s_cand + parity * (CL_MAXC * 4);

            // (1) warp candidate; its owner parks the whole window in shared memory
            const bool live = pos >= j;
            const KT key = live ? K::of(a[c]) : (KT)0;
            const KT wmax = K::warp_max(key);
            const bool cand = live && (key == wmax);
            const unsigned wpos = __reduce_min_sync(kFullMask, cand ? (unsigned)pos : (unsigned)INT_MAX);
            if (cand && (unsigned)pos == wpos) {
#pragma unroll
                for (int v4 = 0; v4 < W / VEC; ++v4) {
                    V16 v;
#pragma unroll
                    for (int e = 0; e < VEC; ++e) v.v[e] = a[v4 * VEC + e];
                    *reinterpret_cast<V16*>(&s_wrow[warp * W + v4 * VEC]) = v;
                }
            }
            if (lane == 0) {
                s_wkey[warp] = wmax;
                s_wpos[warp] = (int)wpos;
            }
            CL_STAMP(0);
            __syncthreads();
            CL_STAMP(1);

            // (2) every warp derives the CTA candidate; warp p pushes it to CTA p of the cluster
            {
                const KT k16 = lane < NW ? s_wkey[lane] : (KT)0;
                const unsigned p16 = lane < NW ? (unsigned)s_wpos[lane] : 0xffffffffu;
                const KT cmax = K::warp_max(k16);
                const bool c16 = (lane < NW) && (k16 == cmax);
                const unsigned cpos = __reduce_min_sync(kFullMask, c16 ? p16 : 0xffffffffu);
                const int bw = __ffs(__ballot_sync(kFullMask, c16 && p16 == cpos)) - 1;
                if (warp < C) {
                    T* dst = cluster.map_shared_rank(rows_p + rank * W, warp);
                    for (int i = lane; i < W; i += 32) dst[i] = s_wrow[bw * W + i];
                    if (lane == 0) {
                        unsigned long long* dc = cluster.map_shared_rank(cand_p + rank * 4, warp);
                        const T pv = s_wrow[bw * W + c];
                        const T rc = (cmax == 0) ? T(0) : T(1) / pv;  // A::one() / pivot (getrf.rs:76)
                        unsigned long long rbits;
                        if (sizeof(T) == 8) rbits = (unsigned long long)__double_as_longlong((double)rc);
                        else rbits = (unsigned long long)__float_as_uint((float)rc);
                        dc[0] = (unsigned long long)cmax;
                        dc[1] = (unsigned long long)cpos;
                        dc[2] = rbits;
                    }
                }
            }
            CL_STAMP(2);
            cluster.sync();
            CL_STAMP(3);

            // (3) every warp picks the same winner from its own shared memory
            const KT gk = lane < C ? (KT)cand_p[lane * 4 + 0] : (KT)0;
            const unsigned gp = lane < C ? (unsigned)cand_p[lane * 4 + 1] : 0xffffffffu;
            const KT gmax = K::warp_max(gk);
            const bool c2 = (lane < C) && (gk == gmax);
            const unsigned gpos_u = __reduce_min_sync(kFullMask, c2 ? gp : 0xffffffffu);
            const int gw = __ffs(__ballot_sync(kFullMask, c2 && gp == gpos_u)) - 1;
            const int gpos = (int)gpos_u;
            const bool sing = (gmax == 0);
            if (rank == 0 && tid == 0) {
                ipiv[j] = row_base + gpos;
                if (sing) *info = step_base + j;  // last zero-pivot step wins (getrf.rs:72-73)
            }
            CL_STAMP(4);
            const bool was_j = (pos == j), was_w = (pos == gpos);
            if (was_j) pos = gpos;
            if (was_w) {
                pos = j;
                // this row is now row j of U: park its live entries (columns j ..)
#pragma unroll
                for (int i = c; i < W; ++i)
                    if (i < rem) my_out[jb + i] = a[i];
            }
            if (pos > j) {
                const T* urow = rows_p + gw * W;
                if (!sing) {
                    const unsigned long long rbits = cand_p[gw * 4 + 2];
                    T recip;
                    if (sizeof(T) == 8) recip = (T)__longlong_as_double((long long)rbits);
                    else recip = (T)__uint_as_float((unsigned)rbits);
                    const T l = a[c] * recip;
                    a[c] = l;
#pragma unroll
                    for (int v4 = (c + 1) / VEC; v4 < W / VEC; ++v4) {
                        if (v4 * VEC < rem) {  // uniform: skip the dead tail of the window
                            const V16 u = *reinterpret_cast<const V16*>(urow + v4 * VEC);
#pragma unroll
                            for (int e = 0; e < VEC; ++e) {
                                const int i = v4 * VEC + e;
                                if (i > c) a[i] -= l * u.v[e];
                            }
                        }
                    }
                }
                my_out[j] = a[c];  // multiplier (or the untouched entry of a singular step)
            }
            CL_STAMP(5);
            if (timing && rank == 0 && tid == 0) g_cl_timing[6] += 1;
        }
        // rotate the window: column jb+GS moves to index 0
#pragma unroll
        for (int i = 0; i < W - GS; ++i) a[i] = a[i + GS];
#pragma unroll
        for (int i = W - GS; i < W; ++i) a[i] = T(0);
    }

    // ---- rows to their final positions (each thread wrote only its own s_out row) ----
    if (pos >= 0) {
        T* p = A + (long long)pos * lda;
        if (vec_ok) {
#pragma unroll
            for (int c = 0; c < W / VEC; ++c) {
                V16 v;
#pragma unroll
                for (int e = 0; e < VEC; ++e) v.v[e] = my_out[c * VEC + e];
                *reinterpret_cast<V16*>(p + c * VEC) = v;
            }
        } else {
            for (int c = 0; c < w; ++c) p[c] = my_out[c];
        }
    }
    cluster.sync();  // no CTA leaves while a peer could still address its shared memory
}

template <class T, int W, int GS>
int launch_cluster(int64_t rows, int64_t w, T* d_a, int64_t lda, int32_t* d_ipiv, int32_t row_base, int32_t* d_info,
                   int32_t step_base, cudaStream_t s) {
    auto kern = panel_cluster_kernel<T, W, GS>;
    const size_t smem = ClusterSmem<T, W>::total;
    static int max_cluster = -1;  // largest cluster size this device accepts for the kernel
    static uint64_t seen_epoch = 0;
    if (stale_for_context(seen_epoch)) max_cluster = -1;
    if (max_cluster < 0) {
        LAIR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        max_cluster = 8;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(16);
            cfg.blockDim = dim3(CL_TPB);
            cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 16;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            int nclusters = 0;
            if (cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg) == cudaSuccess && nclusters >= 1) max_cluster = 16;
        }
        (void)cudaGetLastError();
    }
    int need = (int)((rows + CL_TPB - 1) / CL_TPB);
    int csize = 1;
    while (csize < need) csize *= 2;
    if (csize > max_cluster) return LAIR_B200_ERR_UNSUPPORTED;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(csize);
    cfg.blockDim = dim3(CL_TPB);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csize;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    ProfScope prof(kProfPanel, s, 2.0 * (double)rows * (double)w * sizeof(T));
    LAIR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, d_a, (long long)lda, (int)rows, (int)w, d_ipiv, (int)row_base, d_info,
                                       (int)step_base, (int)ctx().opt.panel_timing));
    LAIR_LAUNCH_CHECK();
    return LAIR_B200_OK;
}

}  // namespace

// Returns LAIR_B200_ERR_UNSUPPORTED (without setting an error) when the panel does not fit one cluster.
template <class T>
int panel_cluster_dev(int64_t rows, int64_t w, T* d_a, int64_t lda, int32_t* d_ipiv, int32_t row_base, int32_t* d_info,
                      int32_t step_base, cudaStream_t s) {
    if (w > 32 || rows > (int64_t)CL_MAXC * CL_TPB) return LAIR_B200_ERR_UNSUPPORTED;
    if (ctx().opt.panel_group == 1) return launch_cluster<T, 32, 1>(rows, w, d_a, lda, d_ipiv, row_base, d_info, step_base, s);
    if (ctx().opt.panel_group == 2) return launch_cluster<T, 32, 2>(rows, w, d_a, lda, d_ipiv, row_base, d_info, step_base, s);
    if (ctx().opt.panel_group == 8) return launch_cluster<T, 32, 8>(rows, w, d_a, lda, d_ipiv, row_base, d_info, step_base, s);
    return launch_cluster<T, 32, 4>(rows, w, d_a, lda, d_ipiv, row_base, d_info, step_base, s);
}

int panel_cluster_max_rows() { return CL_MAXC * CL_TPB; }

// Debug: read (and optionally clear) the phase cycle counters of the cluster panel kernel.
int panel_cluster_timing(long long* out8, bool clear) {
    LAIR_CUDA_CHECK(cudaDeviceSynchronize());
    LAIR_CUDA_CHECK(cudaMemcpyFromSymbol(out8, g_cl_timing, 8 * sizeof(long long)));
    if (clear) {
        long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        LAIR_CUDA_CHECK(cudaMemcpyToSymbol(g_cl_timing, z, sizeof(z)));
    }
    return LAIR_B200_OK;
}

template int panel_cluster_dev<float>(int64_t, int64_t, float*, int64_t, int32_t*, int32_t, int32_t*, int32_t, cudaStream_t);
template int panel_cluster_dev<double>(int64_t, int64_t, double*, int64_t, int32_t*, int32_t, int32_t*, int32_t, cudaStream_t);

}  // namespace lair
