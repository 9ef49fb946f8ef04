// Panel factorization, third generation: one thread-block cluster, the (rows x 32) panel
// resident in SHARED memory, factored as four 8-column sub-panels held in REGISTERS.
//
// Why: measured on B200 (profiles/r1_panel_phases.md) the per-column cost of the earlier
// kernels (one row per thread, 16 warps per CTA) was ~4000 cycles, dominated by fixed
// per-warp work (REDUX arg-max, barriers, serialized shared-memory latency) and only ~3 % by
// the rank-1 update itself.  This kernel keeps the same exchange (candidate rows pushed
// through distributed shared memory, one cluster barrier per column) but:
//   * every thread owns RPT rows, so a CTA of 512 rows needs only 512/RPT threads -- 4x
//     fewer warps contend for the reduction units and barrier skew shrinks;
//   * only the active 8-column sub-panel lives in registers (static indices, 8 column steps
//     compiled once and reused for the four sub-panels -> small instruction footprint);
//     the other columns stay parked in shared memory and are updated ONCE per sub-panel with
//     a rank-8 update (8 FMAs per element per smem round trip instead of 1);
//   * the arg-max uses the top 32 bits of |x| as a coarse key (one REDUX + one vote); the
//     exact 64-bit comparison runs only among lanes that tie on the coarse key.
// In-kernel algorithm per sub-panel s (columns 8s..8s+7), exactly the blocked LU recursion
// of the reference's recursive variant (src/lapack/getrf.rs:216-322) at width 8:
//   8 x { arg-max (src/blas/iamax.rs:6-21) -> exchange -> scale by reciprocal, rank-1 update
//         of the sub-panel columns in registers },
//   U12 = L11^-1 * (pivot rows' parked columns)        [trsm, redundantly per CTA, tiny]
//   parked columns of live rows -= L21 * U12           [rank-8 update from registers]
// Row interchanges stay logical (`pos`), rows are written to their final positions once.
#include <climits>
#include <cooperative_groups.h>

#include "common.cuh"
#include "pivot_key.cuh"

namespace cg = cooperative_groups;

namespace lair {
namespace {

constexpr int PB_W = 32;       // panel width handled by one launch
constexpr int PB_SW = 8;       // sub-panel width (register resident)
constexpr int PB_ROWS = 512;   // rows per CTA
constexpr int PB_MAXC = 16;    // CTAs per cluster

__device__ long long g_pb_timing[8];

template <class T>
struct PBSmem {
    static constexpr int VEC = 16 / sizeof(T);
    static constexpr int LD = PB_W + VEC;  // row pitch: 16-byte aligned rows, conflict-free 128-bit row access
    static constexpr size_t panel_bytes = (size_t)PB_ROWS * LD * sizeof(T);          // the CTA's rows
    static constexpr size_t rows_bytes = (size_t)2 * PB_MAXC * PB_W * sizeof(T);     // pushed candidate rows
    static constexpr size_t cand_bytes = (size_t)2 * PB_MAXC * 4 * sizeof(unsigned long long);
    static constexpr size_t piv_bytes = (size_t)PB_SW * PB_W * sizeof(T);            // the sub-panel's pivot rows
    static constexpr size_t total = panel_bytes + rows_bytes + cand_bytes + piv_bytes + 64;
};

// Exact arg-max of (key, pos) over a warp: larger key wins, ties -> smaller pos.  Coarse pass
// on the top 32 bits; the full comparison only among lanes tying on it.
template <class KT>
__device__ __forceinline__ void warp_argmax(KT key, unsigned pos, KT& kbest, unsigned& pbest) {
    if (sizeof(KT) == 8) {
        const uint32_t hi = (uint32_t)((unsigned long long)key >> 32);
        const uint32_t mh = __reduce_max_sync(kFullMask, hi);
        const unsigned tie = __ballot_sync(kFullMask, hi == mh);
        if (__popc(tie) == 1) {
            const int src = __ffs(tie) - 1;
            kbest = (KT)__shfl_sync(kFullMask, (unsigned long long)key, src);
            pbest = __shfl_sync(kFullMask, pos, src);
            return;
        }
        const uint32_t lo = (hi == mh) ? (uint32_t)key : 0u;
        const uint32_t ml = __reduce_max_sync(kFullMask, lo);
        const bool c = (hi == mh) && ((uint32_t)key == ml);
        kbest = (KT)(((unsigned long long)mh << 32) | ml);
        pbest = __reduce_min_sync(kFullMask, c ? pos : 0xffffffffu);
    } else {
        const uint32_t k32 = (uint32_t)key;
        const uint32_t mk = __reduce_max_sync(kFullMask, k32);
        const unsigned tie = __ballot_sync(kFullMask, k32 == mk);
        kbest = (KT)mk;
        if (__popc(tie) == 1) {
            pbest = __shfl_sync(kFullMask, pos, __ffs(tie) - 1);
            return;
        }
        pbest = __reduce_min_sync(kFullMask, (k32 == mk) ? pos : 0xffffffffu);
    }
}

template <class T, int RPT>
__global__ void __launch_bounds__(PB_ROWS / RPT, 1)
panel_blocked_kernel(T* __restrict__ A, long long lda, int M, int w, int32_t* __restrict__ ipiv, int row_base,
                     int32_t* __restrict__ info, int step_base, int timing) {
    using K = PivotKey<T>;
    using KT = typename K::type;
    using SM = PBSmem<T>;
    constexpr int TPB = PB_ROWS / RPT;
    constexpr int NW = TPB / 32;
    constexpr int VEC = SM::VEC;
    constexpr int LD = SM::LD;
    constexpr int W = PB_W, SW = PB_SW;
    struct alignas(16) V16 { T v[VEC]; };

    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();

    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s_panel = reinterpret_cast<T*>(smem_raw);                                                 // [PB_ROWS][LD]
    T* s_rows = reinterpret_cast<T*>(smem_raw + SM::panel_bytes);                                // [2][MAXC][W]
    unsigned long long* s_cand = reinterpret_cast<unsigned long long*>(smem_raw + SM::panel_bytes + SM::rows_bytes);  // [2][MAXC][4]
    T* s_piv = reinterpret_cast<T*>(smem_raw + SM::panel_bytes + SM::rows_bytes + SM::cand_bytes);  // [SW][W]
    __shared__ KT s_wkey[NW];
    __shared__ unsigned s_wpos[NW];
    __shared__ int s_wrow[NW];  // local row index (0..PB_ROWS) of each warp's candidate

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // debug phase timing: accumulated in registers, flushed once at the end (a global RMW per
    // stamp would add an L2 round trip to every phase it tries to measure)
    long long tprev = 0, tacc0 = 0, tacc1 = 0, tacc2 = 0, tacc3 = 0, tacc4 = 0, tacc5 = 0, tcols = 0;
    const long long tstart = clock64();
#define PB_STAMP(slot)                                  \
    do {                                                \
        if (timing && rank == 0 && tid == 0) {          \
            const long long now_ = clock64();           \
            tacc##slot += now_ - tprev;                 \
            tprev = now_;                               \
        }                                               \
    } while (0)

    // ---- stage this CTA's rows: coalesced global -> shared (row r of the CTA = panel row rank*512 + r) ----
    const int cta_row0 = rank * PB_ROWS;
    const bool vec_ok = (w == W) && ((lda % VEC) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    if (vec_ok) {
        constexpr int CPR = W / VEC;
        for (int c = tid; c < PB_ROWS * CPR; c += TPB) {
            const int r = c / CPR, cc = (c % CPR) * VEC;
            V16 v;
            if (cta_row0 + r < M) v = *reinterpret_cast<const V16*>(A + (long long)(cta_row0 + r) * lda + cc);
            else
#pragma unroll
                for (int e = 0; e < VEC; ++e) v.v[e] = T(0);
            *reinterpret_cast<V16*>(s_panel + r * LD + cc) = v;
        }
    } else {
        for (int idx = tid; idx < PB_ROWS * W; idx += TPB) {
            const int r = idx / W, c = idx % W;
            s_panel[r * LD + c] = (cta_row0 + r < M && c < w) ? A[(long long)(cta_row0 + r) * lda + c] : T(0);
        }
    }
    // thread t owns local rows t, t + TPB, ... (lanes touch consecutive shared-memory rows)
    int pos[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int grow = cta_row0 + tid + r * TPB;
        pos[r] = grow < M ? grow : -1;
    }
    __syncthreads();
    cluster.sync();  // every CTA of the cluster is running before the first remote store

    for (int sb = 0; sb < w; sb += SW) {  // sub-panels
        // ---- sub-panel columns into registers ----
        T a[RPT][SW];
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const T* prow = s_panel + (tid + r * TPB) * LD + sb;
#pragma unroll
            for (int c = 0; c < SW / VEC; ++c) {
                const V16 v = *reinterpret_cast<const V16*>(prow + c * VEC);
#pragma unroll
                for (int e = 0; e < VEC; ++e) a[r][c * VEC + e] = v.v[e];
            }
        }

        // Column steps: a ROLLED loop (the body is compiled once).  The register window slides:
        // a[r][0] is always the current column, a[r][k] column j+k; the rank-1 update writes its
        // result one slot down, so the shift costs nothing.  Finished entries (multipliers, and
        // the U part of a row when it becomes a pivot) go straight to the shared-memory panel.
#pragma unroll 1
        for (int c = 0; c < SW; ++c) {
            const int j = sb + c;
            if (j >= w) break;  // uniform
            const int parity = j & 1;
            const int left = SW - c;  // window entries still inside the sub-panel
            if (timing && rank == 0 && tid == 0) tprev = clock64();
            T* rows_p = s_rows + parity * (PB_MAXC * W);
            unsigned long long* cand_p = s_cand + parity * (PB_MAXC * 4);

            // (1) thread candidate over its live rows, then warp candidate
            KT bkey = 0;
            unsigned bpos = 0x7fffffffu;
            int br = 0;
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const bool live = pos[r] >= j;
                const KT key = live ? K::of(a[r][0]) : (KT)0;
                const unsigned p = live ? (unsigned)pos[r] : 0x7fffffffu;
                if (key > bkey || (key == bkey && p < bpos)) {
                    bkey = key;
                    bpos = p;
                    br = r;
                }
            }
            KT wkey;
            unsigned wpos;
            warp_argmax<KT>(bkey, bpos, wkey, wpos);
            if (bpos == wpos && wpos != 0x7fffffffu) {
                // owner: make its shared-memory row current (columns j .. sb+SW-1 live in the window)
                const int lrow = tid + br * TPB;
                T* prow = s_panel + lrow * LD + j;
#pragma unroll
                for (int r = 0; r < RPT; ++r) {
                    if (r == br) {
#pragma unroll
                        for (int k = 0; k < SW; ++k)
                            if (k < left) prow[k] = a[r][k];
                    }
                }
                s_wrow[warp] = lrow;
            }
            if (lane == 0) {
                s_wkey[warp] = wkey;
                s_wpos[warp] = wpos;
            }
            PB_STAMP(0);
            __syncthreads();
            PB_STAMP(1);

            // (2) CTA candidate (every warp, serial over the few warps), pushed to every peer
            {
                KT ckey = s_wkey[0];
                unsigned cpos = s_wpos[0];
                int cw = 0;
#pragma unroll
                for (int i = 1; i < NW; ++i) {
                    const KT k = s_wkey[i];
                    const unsigned p = s_wpos[i];
                    if (k > ckey || (k == ckey && p < cpos)) {
                        ckey = k;
                        cpos = p;
                        cw = i;
                    }
                }
                const T* src = s_panel + s_wrow[cw] * LD;  // garbage index is never used when the CTA has no live row
                const bool has = (cpos != 0x7fffffffu);
                for (int peer = warp; peer < C; peer += NW) {
                    T* dst = cluster.map_shared_rank(rows_p + rank * W, peer);
                    if (has)
                        for (int i = lane; i < W; i += 32) dst[i] = src[i];
                    if (lane == 0) {
                        unsigned long long* dc = cluster.map_shared_rank(cand_p + rank * 4, peer);
                        T rc = T(0);
                        if (has && ckey != 0) rc = T(1) / src[j];  // A::one() / pivot (getrf.rs:76)
                        unsigned long long rbits;
                        if (sizeof(T) == 8) rbits = (unsigned long long)__double_as_longlong((double)rc);
                        else rbits = (unsigned long long)__float_as_uint((float)rc);
                        dc[0] = (unsigned long long)ckey;
                        dc[1] = (unsigned long long)cpos;
                        dc[2] = rbits;
                    }
                }
            }
            PB_STAMP(2);
            cluster.sync();
            PB_STAMP(3);

            // (3) every warp picks the same winner among the C candidates in its own shared memory
            KT gkey;
            unsigned gpos_u;
            {
                const KT k = lane < C ? (KT)cand_p[lane * 4 + 0] : (KT)0;
                const unsigned p = lane < C ? (unsigned)cand_p[lane * 4 + 1] : 0x7fffffffu;
                warp_argmax<KT>(k, p, gkey, gpos_u);
            }
            // the winning CTA = the lane whose candidate position equals gpos (positions are unique)
            const unsigned mine = lane < C ? (unsigned)cand_p[lane * 4 + 1] : 0xffffffffu;
            const int gw = __ffs(__ballot_sync(kFullMask, mine == gpos_u)) - 1;
            const int gpos = (int)gpos_u;
            const bool sing = (gkey == 0);
            const T* urow = rows_p + gw * W;
            if (rank == 0 && tid == 0) {
                ipiv[j] = row_base + gpos;
                if (sing) *info = step_base + j;  // last zero-pivot step wins (getrf.rs:72-73)
            }
            if (warp == 0) s_piv[c * W + lane] = urow[lane];  // keep the pivot row for the block update (W == 32)
            PB_STAMP(4);
            T recip = T(0);
            if (!sing) {
                const unsigned long long rbits = cand_p[gw * 4 + 2];
                if (sizeof(T) == 8) recip = (T)__longlong_as_double((long long)rbits);
                else recip = (T)__uint_as_float((unsigned)rbits);
            }
            T u[SW];  // u[k] = pivot-row entry of column j+k (zero beyond the sub-panel)
#pragma unroll
            for (int k = 1; k < SW; ++k) u[k] = (k < left) ? urow[j + k] : T(0);
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const bool was_j = (pos[r] == j), was_w = (pos[r] == gpos);
                if (was_j) pos[r] = gpos;
                if (was_w) pos[r] = j;
                if (pos[r] > j) {
                    // live row: multiplier into the panel, update fused with the window shift
                    if (!sing) {
                        const T l = a[r][0] * recip;
                        s_panel[(tid + r * TPB) * LD + j] = l;
#pragma unroll
                        for (int k = 1; k < SW; ++k) a[r][k - 1] = a[r][k] - l * u[k];
                    } else {
                        s_panel[(tid + r * TPB) * LD + j] = a[r][0];
#pragma unroll
                        for (int k = 1; k < SW; ++k) a[r][k - 1] = a[r][k];
                    }
                    a[r][SW - 1] = T(0);
                }
                // rows that are (or were) pivots keep a stale window: their entries already sit in the panel
            }
            PB_STAMP(5);
            if (timing && rank == 0 && tid == 0) tcols += 1;
        }

        const int c1 = sb + SW;           // first parked column
        const int npark = W - c1;         // parked columns still to update (multiple of 8, may be 0)
        const int kdone = (w - sb) < SW ? (w - sb) : SW;  // pivots found in this sub-panel
        if (npark > 0 && c1 < w) {
            __syncthreads();  // s_piv rows complete (written by warp 0 during the column steps)
            // U12 = L11^-1 * P12 by forward substitution, one thread per parked column (tiny)
            if (tid < npark) {
                T uc[SW];
#pragma unroll
                for (int i = 0; i < SW; ++i) uc[i] = s_piv[i * W + c1 + tid];
#pragma unroll
                for (int i = 1; i < SW; ++i) {
                    if (i < kdone) {
#pragma unroll
                        for (int k = 0; k < i; ++k) uc[i] -= s_piv[i * W + sb + k] * uc[k];
                    }
                }
#pragma unroll
                for (int i = 0; i < SW; ++i) s_piv[i * W + c1 + tid] = uc[i];
            }
            __syncthreads();
            // parked columns: pivot rows of this sub-panel take their U12 row; live rows get the rank-8 update
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                T* prow = s_panel + (tid + r * TPB) * LD;
                const int p = pos[r];
                if (p >= sb && p < sb + kdone) {
                    const T* urow12 = s_piv + (p - sb) * W;
                    for (int k = c1; k < W; k += VEC) *reinterpret_cast<V16*>(prow + k) = *reinterpret_cast<const V16*>(urow12 + k);
                } else if (p >= sb + kdone) {
                    T lm[SW];  // this row's 8 multipliers of the sub-panel
#pragma unroll
                    for (int cc = 0; cc < SW / VEC; ++cc) {
                        const V16 v = *reinterpret_cast<const V16*>(prow + sb + cc * VEC);
#pragma unroll
                        for (int e = 0; e < VEC; ++e) lm[cc * VEC + e] = v.v[e];
                    }
                    for (int k = c1; k < W; k += VEC) {
                        V16 x = *reinterpret_cast<const V16*>(prow + k);
#pragma unroll
                        for (int i = 0; i < SW; ++i) {
                            const V16 uu = *reinterpret_cast<const V16*>(s_piv + i * W + k);
#pragma unroll
                            for (int e = 0; e < VEC; ++e) x.v[e] -= lm[i] * uu.v[e];
                        }
                        *reinterpret_cast<V16*>(prow + k) = x;
                    }
                }
            }
            __syncthreads();
        }
    }

    // ---- rows to their final positions: a warp per row, coalesced ----
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        if (pos[r] >= 0) {
            const T* prow = s_panel + (tid + r * TPB) * LD;
            T* g = A + (long long)pos[r] * lda;
            if (vec_ok) {
#pragma unroll
                for (int cc = 0; cc < W / VEC; ++cc) *reinterpret_cast<V16*>(g + cc * VEC) = *reinterpret_cast<const V16*>(prow + cc * VEC);
            } else {
                for (int cc = 0; cc < w; ++cc) g[cc] = prow[cc];
            }
        }
    }
    cluster.sync();  // no CTA leaves while a peer could still address its shared memory
    if (timing && rank == 0 && tid == 0) {
        g_pb_timing[0] += tacc0;
        g_pb_timing[1] += tacc1;
        g_pb_timing[2] += tacc2;
        g_pb_timing[3] += tacc3;
        g_pb_timing[4] += tacc4;
        g_pb_timing[5] += tacc5;
        g_pb_timing[6] += tcols;
        g_pb_timing[7] += clock64() - tstart;  // whole kernel, thread 0 of rank 0
    }
#undef PB_STAMP
}

template <class T, int RPT>
int launch_blocked(int64_t rows, int64_t w, T* d_a, int64_t lda, int32_t* d_ipiv, int32_t row_base, int32_t* d_info,
                   int32_t step_base, cudaStream_t s) {
    auto kern = panel_blocked_kernel<T, RPT>;
    constexpr int TPB = PB_ROWS / RPT;
    const size_t smem = PBSmem<T>::total;
    static int max_cluster = -1;
    if (max_cluster < 0) {
        LAIR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        max_cluster = 8;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(16);
            cfg.blockDim = dim3(TPB);
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
    int need = (int)((rows + PB_ROWS - 1) / PB_ROWS);
    int csize = 1;
    while (csize < need) csize *= 2;
    if (csize > max_cluster) return LAIR_B200_ERR_UNSUPPORTED;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(csize);
    cfg.blockDim = dim3(TPB);
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
int panel_blocked_dev(int64_t rows, int64_t w, T* d_a, int64_t lda, int32_t* d_ipiv, int32_t row_base, int32_t* d_info,
                      int32_t step_base, cudaStream_t s) {
    if (w > PB_W || rows > (int64_t)PB_MAXC * PB_ROWS) return LAIR_B200_ERR_UNSUPPORTED;
    if (ctx().opt.panel_rpt == 2) return launch_blocked<T, 2>(rows, w, d_a, lda, d_ipiv, row_base, d_info, step_base, s);
    if (ctx().opt.panel_rpt == 8) return launch_blocked<T, 8>(rows, w, d_a, lda, d_ipiv, row_base, d_info, step_base, s);
    return launch_blocked<T, 4>(rows, w, d_a, lda, d_ipiv, row_base, d_info, step_base, s);
}

int panel_blocked_timing(long long* out8, bool clear) {
    LAIR_CUDA_CHECK(cudaDeviceSynchronize());
    LAIR_CUDA_CHECK(cudaMemcpyFromSymbol(out8, g_pb_timing, 8 * sizeof(long long)));
    if (clear) {
        long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        LAIR_CUDA_CHECK(cudaMemcpyToSymbol(g_pb_timing, z, sizeof(z)));
    }
    return LAIR_B200_OK;
}

template int panel_blocked_dev<float>(int64_t, int64_t, float*, int64_t, int32_t*, int32_t, int32_t*, int32_t, cudaStream_t);
template int panel_blocked_dev<double>(int64_t, int64_t, double*, int64_t, int32_t*, int32_t, int32_t*, int32_t, cudaStream_t);

}  // namespace lair
