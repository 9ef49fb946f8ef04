// Panel factorization, third generation: one thread-block cluster, the (rows x 32) panel
// resident in SHARED memory, factored as four 8-column sub-panels held in REGISTERS.
//
// Why (measurements in profiles/r1_panel_phases.md, profiles/r1_latbench_b200.jsonl): the
// first panel kernels (one row per thread, 16 warps per CTA) spent ~4000 cycles per column,
// only ~3 % of it in the rank-1 update; every phase was a chain of dependent 20-30-cycle
// instructions, and a DSMEM *push* exchange costs a warp one remote round trip per peer
// (1442 cycles for 16 CTAs) while remote *loads* pipeline.  This kernel therefore:
//   * gives every thread RPT rows, so a CTA of 512 rows is 512/RPT threads -- fewer warps
//     contend for the reduction units and the block barrier is cheap;
//   * keeps only the active 8-column sub-panel in registers, as a SLIDING window (a[r][0] is
//     always the current column; the rank-1 update writes its result one slot down), so the
//     column step is a rolled loop with a small instruction footprint; the other columns stay
//     parked in shared memory and are updated ONCE per sub-panel with a rank-8 update;
//   * exchanges candidates without a cluster barrier in the column loop (default, ASYNC): each CTA
//     pushes its 16-byte {key, position} record to every peer with st.async, the receiver's
//     mbarrier counts the bytes, warp 0 picks the winner locally and PULLS only the winner's row
//     (remote loads pipeline).  The earlier exchange (publish locally, ONE cluster barrier, every
//     warp pulls records and rows of its share of the peers) is kept as ASYNC = false;
//   * finds arg-max with the top 32 bits of |x| as a coarse key (one REDUX + one vote); the
//     exact 64-bit comparison runs only among lanes that tie on the coarse key.
// In-kernel algorithm per sub-panel s (columns 8s..8s+7) = the blocked LU recursion of the
// reference's recursive variant (src/lapack/getrf.rs:216-322) at width 8:
//   8 x { arg-max (src/blas/iamax.rs:6-21) -> exchange -> scale by reciprocal, rank-1 update
//         of the sub-panel columns in registers (src/lapack/getrf.rs:76-87) },
//   U12 = L11^-1 * (pivot rows' parked columns)        [trsm, redundantly per CTA, tiny]
//   parked columns of live rows -= L21 * U12           [rank-8 update]
// Row interchanges stay logical (`pos`); rows are written to their final positions once.
#include <climits>
#include <cooperative_groups.h>

#include "common.cuh"
#include "pivot_key.cuh"

namespace cg = cooperative_groups;

namespace lair {
namespace {

constexpr int PB_W = 32;       // default panel width handled by one launch (a 64-wide variant exists)
constexpr int PB_SW = 8;       // sub-panel width (register resident)
constexpr int PB_ROWS32 = 512; // rows per CTA, 32-wide panels
template <class T> constexpr int pb_rows64() { return sizeof(T) == 8 ? 256 : 512; }  // rows per CTA, 64-wide panels
constexpr int PB_MAXC = 16;    // CTAs per cluster
constexpr unsigned PB_NOPOS = 0x7fffffffu;

__device__ long long g_pb_timing[8];

template <class T, int W, int ROWS>
struct PBSmem {
    static constexpr int VEC = 16 / sizeof(T);
    static constexpr int LD = W + VEC;  // row pitch: 16-byte aligned rows, conflict-free 128-bit row access
    static constexpr size_t panel_bytes = (size_t)ROWS * LD * sizeof(T);           // the CTA's rows
    static constexpr size_t rows_bytes = (size_t)PB_MAXC * W * sizeof(T);          // pulled candidate rows
    static constexpr size_t mine_bytes = (size_t)2 * W * sizeof(T);                // published candidate row (2 parities)
    static constexpr size_t piv_bytes = (size_t)PB_SW * W * sizeof(T);             // the sub-panel's pivot rows
    static constexpr size_t total = panel_bytes + rows_bytes + mine_bytes + piv_bytes + 64;
};

// warp_argmax (exact, coarse-key first) lives in pivot_key.cuh.

// ---- st.async exchange helpers (ASYNC variant) ----
__device__ __forceinline__ unsigned pb_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned pb_mapa(unsigned addr, unsigned cta_rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void pb_mbar_wait_cluster(unsigned bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}

// ASYNC = false: candidates published locally, one cluster barrier per column, peers PULL records
//                and rows (second generation of the exchange).
// ASYNC = true : every CTA PUSHES its 16-byte candidate record to all peers with st.async
//                (completion counted by the receiver's mbarrier: no cluster barrier in the column
//                loop; 448 vs ~1900 cycles for 16 CTAs, profiles/r1_latbench_b200.jsonl), each warp
//                then decides the winner from its own shared memory and pulls only the winner's row.
// W x PB_ROWS_ per CTA: 32 x 512 (any exchange) or 64 x 256 (f64) / 64 x 512 (f32), ASYNC only: a
// whole 64-column block of the outer sweep in ONE launch when its rows fit 16 CTAs.
template <class T, int RPT, bool ASYNC, int W, int PB_ROWS>
__global__ void __launch_bounds__(PB_ROWS / RPT, 1)
panel_blocked_kernel(T* __restrict__ A, long long lda, int M, int w, int32_t* __restrict__ ipiv, int row_base,
                     int32_t* __restrict__ info, int step_base, int timing) {
    using K = PivotKey<T>;
    using KT = typename K::type;
    using SM = PBSmem<T, W, PB_ROWS>;
    constexpr int TPB = PB_ROWS / RPT;
    constexpr int NW = TPB / 32;
    constexpr int VEC = SM::VEC;
    constexpr int LD = SM::LD;
    constexpr int SW = PB_SW;
    constexpr int WL = W / 32;  // panel columns per lane in the row copies
    struct alignas(16) V16 { T v[VEC]; };
    static_assert(W == 32 || (W == 64 && ASYNC), "64-wide panels use the st.async exchange");

    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();

    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s_panel = reinterpret_cast<T*>(smem_raw);                                                  // [PB_ROWS][LD]
    T* s_rows = reinterpret_cast<T*>(smem_raw + SM::panel_bytes);                                 // [MAXC][W] pulled rows
    T* s_mine = reinterpret_cast<T*>(smem_raw + SM::panel_bytes + SM::rows_bytes);                // [2][W] published row
    T* s_piv = reinterpret_cast<T*>(smem_raw + SM::panel_bytes + SM::rows_bytes + SM::mine_bytes);  // [SW][W]
    __shared__ __align__(16) unsigned long long s_pub[2][2];   // published {key, pos} per parity (read remotely)
    __shared__ __align__(16) unsigned long long s_cand[2][PB_MAXC][2];  // pulled (parity 0 only) or pushed {key, pos}
    __shared__ __align__(8) unsigned long long s_mbar[2];               // ASYNC: one mbarrier per column parity
    __shared__ __align__(16) unsigned long long s_win[2];               // ASYNC: the column's winner {key, pos}
    __shared__ T s_recip[PB_MAXC];
    __shared__ KT s_wkey[NW];
    __shared__ unsigned s_wpos[NW];
    __shared__ int s_wrow[NW];  // local row index (0..PB_ROWS) of each warp's candidate
    __shared__ int s_pos[PB_ROWS];  // final position of every local row (write-out)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // debug phase timing: accumulated in registers, flushed once at the end
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
        // cp.async: every 16-byte chunk of the CTA's 128 KB slab is in flight at once (zero-filled past row M)
        constexpr int CPR = W / VEC;
        for (int c = tid; c < PB_ROWS * CPR; c += TPB) {
            const int r = c / CPR, cc = (c % CPR) * VEC;
            const bool in = cta_row0 + r < M;
            const T* src = A + (long long)(in ? cta_row0 + r : 0) * lda + cc;
            const unsigned dst = (unsigned)__cvta_generic_to_shared(s_panel + r * LD + cc);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(in ? 16 : 0) : "memory");
        }
        asm volatile("cp.async.wait_all;\n" ::: "memory");
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
    if (tid < NW) s_wrow[tid] = 0;  // always a valid local row, even before a warp has had a live candidate
    if (ASYNC && tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pb_smem_u32(&s_mbar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pb_smem_u32(&s_mbar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster.sync();  // every CTA of the cluster is running before the first remote access

    for (int sb = 0; sb < w; sb += SW) {  // sub-panels
        // ---- sub-panel columns into the register window ----
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

        // Column steps: a ROLLED loop (the body is compiled once).  a[r][0] is the current
        // column, a[r][k] column j+k; the update writes one slot down, so the shift is free.
#pragma unroll 1
        for (int c = 0; c < SW; ++c) {
            const int j = sb + c;
            if (j >= w) break;  // uniform
            const int parity = j & 1;
            const int left = SW - c;  // window entries still inside the sub-panel
            if (timing && rank == 0 && tid == 0) tprev = clock64();

            // (1) thread candidate over its live rows (keys first, then a compare tree), warp candidate
            KT key[RPT];
            unsigned kp[RPT];
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const bool live = pos[r] >= j;
                key[r] = live ? K::of(a[r][0]) : (KT)0;
                kp[r] = live ? (unsigned)pos[r] : PB_NOPOS;
            }
            int br = 0;
            KT bkey = key[0];
            unsigned bpos = kp[0];
#pragma unroll
            for (int r = 1; r < RPT; ++r) {
                const bool better = key[r] > bkey || (key[r] == bkey && kp[r] < bpos);
                bkey = better ? key[r] : bkey;
                bpos = better ? kp[r] : bpos;
                br = better ? r : br;
            }
            KT wkey;
            unsigned wpos;
            int wsrc;
            warp_argmax<KT>(bkey, bpos, wkey, wpos, wsrc);
            if (lane == wsrc) {
                if (wpos != PB_NOPOS) {
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
                s_wkey[warp] = wkey;
                s_wpos[warp] = wpos;
            }
            PB_STAMP(0);
            __syncthreads();
            PB_STAMP(1);

            // (2) warp 0: CTA candidate -> published in this CTA's own shared memory
            if (warp == 0) {
                const KT k = lane < NW ? s_wkey[lane] : (KT)0;
                const unsigned p = lane < NW ? s_wpos[lane] : PB_NOPOS;
                KT ckey;
                unsigned cpos;
                int cw;
                warp_argmax<KT>(k, p, ckey, cpos, cw);
                const int lrow = s_wrow[cw < NW ? cw : 0];
#pragma unroll
                for (int q = 0; q < WL; ++q)  // garbage when the CTA has no live row: never selected
                    s_mine[parity * W + lane + 32 * q] = s_panel[lrow * LD + lane + 32 * q];
                if constexpr (ASYNC) {
                    const unsigned bar = pb_smem_u32(&s_mbar[parity]);
                    if (C == 1) {  // a single CTA: this warp is also the only consumer, no exchange
                        if (lane == 0) {
                            s_cand[parity][0][0] = (unsigned long long)ckey;
                            s_cand[parity][0][1] = (unsigned long long)cpos;
                        }
                    } else if (lane == 0) {  // arm this column's phase: C records of 16 bytes will land here
                        unsigned long long st_;
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 %0, [%1], %2;" : "=l"(st_) : "r"(bar), "r"((unsigned)C * 16u) : "memory");
                    }
                    __syncwarp();  // the row above is written before any peer can learn of the record
                    if (C > 1 && lane < C) {
                        const unsigned raddr = pb_mapa(pb_smem_u32(&s_cand[parity][rank][0]), (unsigned)lane);
                        const unsigned rbar = pb_mapa(bar, (unsigned)lane);
                        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];" ::"r"(raddr),
                                     "l"((unsigned long long)ckey), "l"((unsigned long long)cpos), "r"(rbar)
                                     : "memory");
                    }
                } else if (lane == 0) {
                    s_pub[parity][0] = (unsigned long long)ckey;
                    s_pub[parity][1] = (unsigned long long)cpos;
                }
            }
            PB_STAMP(2);
            KT gkey;
            unsigned gpos_u;
            int gw;
            const T* urow;
            T recip_w;  // 1 / pivot (getrf.rs:76)
            if constexpr (ASYNC) {
                // (3a) warp 0 waits for the C records, decides the winner and pulls the winner's row: ONE
                //      remote request per CTA (128 warps pulling from one SM cost ~1200 cycles, measured)
                if (warp == 0) {
                    if (C > 1) pb_mbar_wait_cluster(pb_smem_u32(&s_mbar[parity]), (unsigned)(j >> 1) & 1u);
                    PB_STAMP(3);
                    const KT k = lane < C ? (KT)s_cand[parity][lane][0] : (KT)0;
                    const unsigned p = lane < C ? (unsigned)s_cand[parity][lane][1] : PB_NOPOS;
                    warp_argmax<KT>(k, p, gkey, gpos_u, gw);
                    const T* remote = cluster.map_shared_rank(s_mine + parity * W, gw < C ? gw : rank);
                    T v[WL];
#pragma unroll
                    for (int q = 0; q < WL; ++q) v[q] = remote[lane + 32 * q];
#pragma unroll
                    for (int q = 0; q < WL; ++q) {
                        s_rows[lane + 32 * q] = v[q];
                        s_piv[c * W + lane + 32 * q] = v[q];  // keep the pivot row for the block update
                    }
                    if (lane == 0) {
                        s_win[0] = (unsigned long long)gkey;
                        s_win[1] = (unsigned long long)gpos_u;
                    }
                }
                __syncthreads();
                gkey = (KT)s_win[0];
                gpos_u = (unsigned)s_win[1];
                urow = s_rows;
                recip_w = T(1) / urow[j];
            } else {
            cluster.sync();
            PB_STAMP(3);

            // (3) pull: warp wq fetches the record and the row of peers wq, wq+NW, ... (all loads in flight
            //     together); the lane that holds column j also forms the reciprocal of that candidate
            {
                constexpr int Q = (PB_MAXC + NW - 1) / NW;
                T v[Q];
                unsigned long long rec[Q];
#pragma unroll
                for (int q = 0; q < Q; ++q) {  // unconditional loads (own CTA stands in for absent peers)
                    const int peer = warp + q * NW;
                    const int pc = peer < C ? peer : rank;
                    v[q] = cluster.map_shared_rank(s_mine + parity * W, pc)[lane];
                    rec[q] = cluster.map_shared_rank(&s_pub[parity][0], pc)[lane & 1];
                }
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const int peer = warp + q * NW;
                    const T rc = T(1) / v[q];  // A::one() / pivot (getrf.rs:76); only lane j's value is kept
                    if (peer < C) {
                        s_rows[peer * W + lane] = v[q];
                        if (lane < 2) s_cand[0][peer][lane] = rec[q];
                        if (lane == j) s_recip[peer] = rc;
                    }
                }
            }
            __syncthreads();
            // every warp picks the same winner among the C candidates now in its own shared memory
            {
                const KT k = lane < C ? (KT)s_cand[0][lane][0] : (KT)0;
                const unsigned p = lane < C ? (unsigned)s_cand[0][lane][1] : PB_NOPOS;
                warp_argmax<KT>(k, p, gkey, gpos_u, gw);
            }
            urow = s_rows + gw * W;
            if (warp == NW - 1) s_piv[c * W + lane] = urow[lane];  // keep the pivot row for the block update
            recip_w = s_recip[gw];
            }
            const int gpos = (int)gpos_u;
            const bool sing = (gkey == 0);
            if (rank == 0 && tid == 0) {
                ipiv[j] = row_base + gpos;
                if (sing) *info = step_base + j;  // last zero-pivot step wins (getrf.rs:72-73)
            }
            PB_STAMP(4);
            const T recip = sing ? T(0) : recip_w;
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
        if (npark > 0 && c1 < w) {        // (then all SW pivots of this sub-panel exist)
            __syncthreads();  // s_piv rows complete
            // U12 = L11^-1 * P12 by forward substitution, one thread per parked column (tiny)
            if (tid < npark) {
                T uc[SW];
#pragma unroll
                for (int i = 0; i < SW; ++i) uc[i] = s_piv[i * W + c1 + tid];
#pragma unroll
                for (int i = 1; i < SW; ++i) {
#pragma unroll
                    for (int k = 0; k < i; ++k) uc[i] -= s_piv[i * W + sb + k] * uc[k];
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
                if (p >= sb && p < sb + SW) {
                    const T* urow12 = s_piv + (p - sb) * W;
                    for (int k = c1; k < W; k += VEC) *reinterpret_cast<V16*>(prow + k) = *reinterpret_cast<const V16*>(urow12 + k);
                } else if (p >= sb + SW) {
                    T lm[SW];  // this row's 8 multipliers of the sub-panel
#pragma unroll
                    for (int cc = 0; cc < SW / VEC; ++cc) {
                        const V16 v = *reinterpret_cast<const V16*>(prow + sb + cc * VEC);
#pragma unroll
                        for (int e = 0; e < VEC; ++e) lm[cc * VEC + e] = v.v[e];
                    }
#pragma unroll 2
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

    // ---- rows to their final positions ----
    if (vec_ok) {
        // coalesced: the final position of every local row goes through shared memory, then
        // consecutive threads write consecutive 16-byte chunks of the same row (a thread writing its
        // own row chunk by chunk makes every store instruction touch 32 different lines)
#pragma unroll
        for (int r = 0; r < RPT; ++r) s_pos[tid + r * TPB] = pos[r];
        __syncthreads();
        constexpr int CPR = W / VEC;
        for (int c = tid; c < PB_ROWS * CPR; c += TPB) {
            const int r = c / CPR, cc = (c % CPR) * VEC;
            const int p = s_pos[r];
            if (p >= 0) *reinterpret_cast<V16*>(A + (long long)p * lda + cc) = *reinterpret_cast<const V16*>(s_panel + r * LD + cc);
        }
    } else {
        __syncthreads();
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            if (pos[r] >= 0) {
                const T* prow = s_panel + (tid + r * TPB) * LD;
                T* g = A + (long long)pos[r] * lda;
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

template <class T, int RPT, bool ASYNC, int W, int ROWS>
int launch_blocked(int64_t rows, int64_t w, T* d_a, int64_t lda, int32_t* d_ipiv, int32_t row_base, int32_t* d_info,
                   int32_t step_base, cudaStream_t s) {
    auto kern = panel_blocked_kernel<T, RPT, ASYNC, W, ROWS>;
    constexpr int TPB = ROWS / RPT;
    const size_t smem = PBSmem<T, W, ROWS>::total;
    static int max_cluster = -1;
    static uint64_t seen_epoch = 0;
    if (stale_for_context(seen_epoch)) max_cluster = -1;
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
    int need = (int)((rows + ROWS - 1) / ROWS);
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

// Widest panel one launch of this kernel takes for `rows` rows: 64 (option panel_w64, st.async
// exchange) while the rows fit 16 CTAs of the 64-wide layout, else 32, else 0.
template <class T>
int panel_blocked_max_width(int64_t rows) {
    if (ctx().opt.panel_w64 != 0 && ctx().opt.panel_exchange != 0 && rows <= (int64_t)PB_MAXC * pb_rows64<T>()) return 64;
    if (rows <= (int64_t)PB_MAXC * PB_ROWS32) return 32;
    return 0;
}

// Returns LAIR_B200_ERR_UNSUPPORTED (without setting an error) when the panel does not fit one cluster.
template <class T>
int panel_blocked_dev(int64_t rows, int64_t w, T* d_a, int64_t lda, int32_t* d_ipiv, int32_t row_base, int32_t* d_info,
                      int32_t step_base, cudaStream_t s) {
    if (w > 32) {
        if (w > panel_blocked_max_width<T>(rows)) return LAIR_B200_ERR_UNSUPPORTED;
        return launch_blocked<T, 2, true, 64, pb_rows64<T>()>(rows, w, d_a, lda, d_ipiv, row_base, d_info, step_base, s);
    }
    if (rows > (int64_t)PB_MAXC * PB_ROWS32) return LAIR_B200_ERR_UNSUPPORTED;
    const bool async = ctx().opt.panel_exchange != 0;
#define PB_GO(RPT)                                                                                                           \
    return async ? launch_blocked<T, RPT, true, 32, PB_ROWS32>(rows, w, d_a, lda, d_ipiv, row_base, d_info, step_base, s)    \
                 : launch_blocked<T, RPT, false, 32, PB_ROWS32>(rows, w, d_a, lda, d_ipiv, row_base, d_info, step_base, s)
    if (ctx().opt.panel_rpt == 1) PB_GO(1);
    if (ctx().opt.panel_rpt == 4) PB_GO(4);
    PB_GO(2);
#undef PB_GO
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
template int panel_blocked_max_width<float>(int64_t);
template int panel_blocked_max_width<double>(int64_t);

}  // namespace lair
