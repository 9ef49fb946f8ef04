// Panel factorization, fourth generation: the per-column dependent chain cut to
//   warp candidate -> block barrier -> warp 0 pushes {header, register window} to every CTA
//   -> every warp waits for the C records on an mbarrier and decides the winner itself -> update.
//
// Measured background (profiles/r2p_*): the third generation (panel_blocked.cu) spends 2 070-2 370
// cycles per column whatever the cluster size (1 024 rows on 2 CTAs: 2 073; 8 192 rows on 16 CTAs:
// 2 266) -- the time is a chain of ~500 dependent instructions per column, not the exchange.  This
// kernel keeps the algorithm and removes links of that chain:
//   * the winner's row is no longer PULLED over DSMEM after the decision (a remote round trip, a
//     shared-memory hop and a block barrier): the 16-byte record every CTA pushes with st.async
//     carries the candidate row's 8-slot REGISTER WINDOW with it (80 bytes f64 / 48 bytes f32 per
//     peer), so when the records have landed every warp holds the pivot row's active columns;
//   * every warp waits on the mbarrier and repeats the decision over the C record headers, which
//     carry the coarse key ready-made: no "warp 0 decides, block barrier, everybody reads the
//     result" hop;
//   * each of the three arg-max levels (warp, CTA, cluster) is ONE REDUX on the coarse key with the
//     lane number packed into its low five bits; the vote that detects a coarse tie overlaps the
//     speculative loads for the fast answer and only a tie takes the exact path (pp_argmax_hdr);
//   * 1 / pivot (getrf.rs:76) is formed by warp 1 for all warp candidates at once while warp 0
//     reduces and pushes the window chunks, and travels in the header chunk, which goes last;
//   * only the window chunks that still hold live columns travel;
//   * mbarrier waits use the CTA-scope acquire (the data lands in this CTA's own shared memory; the
//     cluster-scope form costs an L1 invalidate per wait);
//   * the pusher's remote addresses are computed once per launch, every lane owns fixed
//     (peer, chunk) pairs;
//   * the rest of a pivot row (its multipliers and the parked columns, needed only at the end of
//     the sub-panel) is pushed by the row's owner CTA one column later, by the warps that idle
//     while warp 0 exchanges, counted on a second mbarrier that is waited for once per sub-panel.
// Algorithm and rounding are unchanged (same operations in the same order as panel_blocked.cu:
// the blocked recursion of src/lapack/getrf.rs:216-322 at width 8, arg-max per src/blas/iamax.rs:6-21,
// scale by the reciprocal and rank-1 update per getrf.rs:76-87), so pivots and L\U are bit-identical
// to the third generation (tools/r2_probe_panel_push.py, tests/test_gpu_parity.py).
#include <climits>
#include <cooperative_groups.h>

#include "common.cuh"
#include "pivot_key.cuh"

namespace cg = cooperative_groups;

namespace lair {
namespace {

constexpr int PP_SW = 8;     // sub-panel width = register window
constexpr int PP_MAXC = 16;  // CTAs per cluster

__device__ long long g_pp_timing[8];

template <class T, int W, int ROWS>
struct PPSmem {
    static constexpr int VEC = 16 / sizeof(T);
    static constexpr int LD = W + VEC;                       // 16-byte aligned rows, conflict-free 128-bit row access
    static constexpr int REC = 16 + PP_SW * (int)sizeof(T);  // record: 16-byte header + the window
    static constexpr size_t panel_bytes = (size_t)ROWS * LD * sizeof(T);
    static constexpr size_t piv_bytes = (size_t)PP_SW * W * sizeof(T);
    static constexpr size_t cand_bytes = (size_t)2 * PP_MAXC * REC;
    static constexpr size_t total = panel_bytes + piv_bytes + cand_bytes + 64;
};

__device__ __forceinline__ unsigned pp_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned pp_mapa(unsigned addr, unsigned cta_rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta_rank));
    return r;
}
// The data an mbarrier phase covers lands in THIS CTA's shared memory (st.async complete_tx), so the default CTA-scope acquire
// orders it; the .acquire.cluster form costs an L1 invalidate (CCTL.IVALL: 18 % of this kernel's stall samples, ncu r2p) per wait.
__device__ __forceinline__ void pp_mbar_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void pp_expect_tx(unsigned bar, unsigned bytes) {
    unsigned long long st_;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 %0, [%1], %2;" : "=l"(st_) : "r"(bar), "r"(bytes) : "memory");
}
// 16 bytes to a (possibly remote) CTA's shared memory, completion counted in bytes by that CTA's mbarrier
__device__ __forceinline__ void pp_push16(unsigned raddr, ulonglong2 v, unsigned rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];" ::"r"(raddr), "l"(v.x), "l"(v.y),
                 "r"(rbar)
                 : "memory");
}

__device__ __forceinline__ double pp_fnma(double l, double u, double a) { return fma(-l, u, a); }
__device__ __forceinline__ float pp_fnma(float l, float u, float a) { return fmaf(-l, u, a); }

// Lane holding the arg-max of (key, pos): larger key wins, ties -> smaller pos (every lane gets the
// same answer).  One REDUX on the top 32 bits + one vote decide almost every call.
template <class KT>
__device__ __forceinline__ int pp_argmax_lane(KT key, unsigned pos) {
    const unsigned lane_bit = 1u << (threadIdx.x & 31);
    const uint32_t hi = (sizeof(KT) == 8) ? (uint32_t)((unsigned long long)key >> 32) : (uint32_t)key;
    const uint32_t mh = __reduce_max_sync(kFullMask, hi);
    unsigned tie = __ballot_sync(kFullMask, hi == mh);
    if (__popc(tie) != 1) {
        if (sizeof(KT) == 8) {
            const uint32_t lo = (hi == mh) ? (uint32_t)key : 0u;
            const uint32_t ml = __reduce_max_sync(kFullMask, lo);
            tie = __ballot_sync(kFullMask, (hi == mh) && ((uint32_t)key == ml));
        }
        if (__popc(tie) != 1) {
            const unsigned pm = __reduce_min_sync(kFullMask, (tie & lane_bit) ? pos : 0xffffffffu);
            tie = __ballot_sync(kFullMask, ((tie & lane_bit) != 0) && pos == pm);
        }
    }
    return __ffs(tie) - 1;
}

// Record: 16-byte header {1 / win[0] (8 bytes, f32 in the low word), coarse key (1 + the top 32 bits of the key of win[0];
// 0 = no live row), posrow = position << 12 | local row} + the candidate row's register window.
constexpr unsigned PP_NOPOSROW = 0xffffffffu;
template <class T> struct PPRc;
template <> struct PPRc<double> {
    __device__ static __forceinline__ unsigned long long bits(double rc) { return (unsigned long long)__double_as_longlong(rc); }
    __device__ static __forceinline__ double from(unsigned long long b) { return __longlong_as_double((long long)b); }
};
template <> struct PPRc<float> {
    __device__ static __forceinline__ unsigned long long bits(float rc) { return (unsigned long long)__float_as_uint(rc); }
    __device__ static __forceinline__ float from(unsigned long long b) { return __uint_as_float((unsigned)b); }
};

// Arg-max lane over record headers: coarse key + posrow decide almost always; `full()` (the exact key) is evaluated only when
// two lanes tie on the coarse key.
template <class KT, class F>
__device__ __forceinline__ int pp_argmax_hdr(uint32_t hi, unsigned posrow, F full) {
    const uint32_t mh = __reduce_max_sync(kFullMask, hi);
    if (mh == 0) return 0;  // no live row behind any lane (live rows carry a coarse key >= 1)
    unsigned tie = __ballot_sync(kFullMask, hi == mh);
    if (tie & (tie - 1)) {
        const unsigned lane_bit = 1u << (threadIdx.x & 31);
        if (sizeof(KT) == 8) {
            const KT key = full();
            const uint32_t lo = (hi == mh) ? (uint32_t)key : 0u;
            const uint32_t ml = __reduce_max_sync(kFullMask, lo);
            tie = __ballot_sync(kFullMask, (hi == mh) && ((uint32_t)key == ml));
        }
        if (tie & (tie - 1)) {
            const unsigned pm = __reduce_min_sync(kFullMask, (tie & lane_bit) ? posrow : 0xffffffffu);
            tie = __ballot_sync(kFullMask, ((tie & lane_bit) != 0) && posrow == pm);
        }
    }
    return __ffs(tie) - 1;
}

// Fast path of the arg-max: one REDUX on the coarse key with the lane number packed into its low five bits names the winning
// lane at once (no vote, no find-first on the dependent chain).  `exact` comes back true when that answer may be wrong -- another
// lane shares the truncated key, or no lane is live -- and the caller then repeats the choice with pp_argmax_hdr; the vote
// that decides this overlaps with whatever the caller issues speculatively for the fast answer.
__device__ __forceinline__ int pp_argmax_fast(uint32_t hi, bool& exact) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t pk = (hi & ~31u) | (31u - lane);
    const uint32_t mp = __reduce_max_sync(kFullMask, pk);
    const int src = 31 - (int)(mp & 31u);
    exact = ((mp >> 5) == 0u) | (__any_sync(kFullMask, ((pk >> 5) == (mp >> 5)) && (int)lane != src) != 0);
    return src;
}

// One cluster factors the (M x w) panel, w <= W; CTA `rank` holds rows rank*ROWS .. +ROWS in shared memory.
template <class T, int RPT, int W, int ROWS>
__global__ void __launch_bounds__(ROWS / RPT, 1)
panel_push_kernel(T* __restrict__ A, long long lda, int M, int w, int32_t* __restrict__ ipiv, int row_base, int32_t* __restrict__ info,
                  int step_base, int timing) {
    using K = PivotKey<T>;
    using KT = typename K::type;
    using SM = PPSmem<T, W, ROWS>;
    using RC = PPRc<T>;
    constexpr int TPB = ROWS / RPT;
    constexpr int NW = TPB / 32;
    constexpr int VEC = SM::VEC;
    constexpr int LD = SM::LD;
    constexpr int SW = PP_SW;
    constexpr int REC = SM::REC;
    constexpr int NCH = REC / 16;  // 16-byte chunks per record (header first)
    struct alignas(16) V16 { T v[VEC]; };
    static_assert(NW >= 3 && NW <= 32, "warps per CTA");
    static_assert(ROWS <= 4096, "posrow keeps the local row in 12 bits");

    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();

    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s_panel = reinterpret_cast<T*>(smem_raw);                                         // [ROWS][LD]
    T* s_piv = reinterpret_cast<T*>(smem_raw + SM::panel_bytes);                         // [SW][W] pivot rows of the sub-panel
    unsigned char* s_cand = smem_raw + SM::panel_bytes + SM::piv_bytes;                  // [2][MAXC] records pushed by the CTAs
    __shared__ __align__(16) unsigned char s_wc[NW * REC];                               // per-warp candidate records
    __shared__ __align__(8) unsigned long long s_mbar[2];                                // one mbarrier per column parity
    __shared__ __align__(8) unsigned long long s_pbar[2];                                // rest of the pivot rows, per sub-panel parity
    __shared__ int s_pos[ROWS];                                                          // final position of every local row (write-out)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long tprev = 0, tacc0 = 0, tacc1 = 0, tacc2 = 0, tacc3 = 0, tacc4 = 0, tacc5 = 0, tcols = 0;
    const long long tstart = clock64();
#define PP_STAMP(slot)                                  \
    do {                                                \
        if (timing == 1 && rank == 0 && tid == 0) {     \
            const long long now_ = clock64();           \
            tacc##slot += now_ - tprev;                 \
            tprev = now_;                               \
        }                                               \
    } while (0)
    // timing == 2: the parts outside the column loop (0 stage-in, 1 window load, 2 column loops, 3 last row's rest + wait,
    // 4 U12 solve, 5 rank-8 update, 6 sub-panels; write-out = total - sum)
#define PP_FINE(slot)                                   \
    do {                                                \
        if (timing == 3 && rank == 0 && tid == 0) {     \
            const long long now_ = clock64();           \
            tacc##slot += now_ - tprev;                 \
            tprev = now_;                               \
        }                                               \
    } while (0)
#define PP_COARSE(slot)                                 \
    do {                                                \
        if (timing == 2 && rank == 0 && tid == 0) {     \
            const long long now_ = clock64();           \
            tacc##slot += now_ - tprev;                 \
            tprev = now_;                               \
        }                                               \
    } while (0)

    // ---- stage this CTA's rows: coalesced global -> shared ----
    const int cta_row0 = rank * ROWS;
    const bool vec_ok = (w == W) && ((lda % VEC) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    if (vec_ok) {
        constexpr int CPR = W / VEC;
        for (int c = tid; c < ROWS * CPR; c += TPB) {
            const int r = c / CPR, cc = (c % CPR) * VEC;
            const bool in = cta_row0 + r < M;
            const T* src = A + (long long)(in ? cta_row0 + r : 0) * lda + cc;
            const unsigned dst = pp_smem_u32(s_panel + r * LD + cc);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(in ? 16 : 0) : "memory");
        }
        asm volatile("cp.async.wait_all;\n" ::: "memory");
    } else {
        for (int idx = tid; idx < ROWS * W; idx += TPB) {
            const int r = idx / W, c = idx % W;
            s_panel[r * LD + c] = (cta_row0 + r < M && c < w) ? A[(long long)(cta_row0 + r) * lda + c] : T(0);
        }
    }
    int pos[RPT];        // thread t owns local rows t, t + TPB, ...: logical position (-1: padding row)
    unsigned kpr[RPT];   // position << 12 | local row while the row is live, PP_NOPOSROW afterwards
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int grow = cta_row0 + tid + r * TPB;
        pos[r] = grow < M ? grow : -1;
        kpr[r] = grow < M ? ((unsigned)grow << 12) | (unsigned)(tid + r * TPB) : PP_NOPOSROW;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pp_smem_u32(&s_mbar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pp_smem_u32(&s_mbar[1])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pp_smem_u32(&s_pbar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pp_smem_u32(&s_pbar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // warp 0's fixed share of the record push: lane -> peer lane & 15, chunks (lane >> 4), +2, +4
    const int my_peer = lane & (PP_MAXC - 1);
    const int my_ch0 = lane >> 4;
    const unsigned push_dst0 = pp_mapa(pp_smem_u32(s_cand + rank * REC), (unsigned)(my_peer < C ? my_peer : 0));
    const unsigned push_bar0 = pp_mapa(pp_smem_u32(&s_mbar[0]), (unsigned)(my_peer < C ? my_peer : 0));
    __syncthreads();
    cluster.sync();  // every CTA of the cluster is running (and its barriers exist) before the first remote access
    if (timing == 2 && rank == 0 && tid == 0) {
        tprev = clock64();
        tacc0 = tprev - tstart;
    }

    int pend_c = -1, pend_lrow = 0;  // CTA-uniform: this CTA owns the pivot row of column pend_c and still has to push its rest

    for (int sb = 0; sb < w; sb += SW) {  // sub-panels
        const int c1 = sb + SW;                       // first parked column
        const int npark = (c1 < w) ? W - c1 : 0;      // parked columns this sub-panel updates (then all SW pivots exist)
        const int sp = (sb >> 3) & 1;
        const int nrest = (W - sb) / VEC;             // 16-byte chunks of a pivot row from column sb on
        if (npark > 0 && tid == 0) pp_expect_tx(pp_smem_u32(&s_pbar[sp]), (unsigned)(SW * (W - sb) * sizeof(T)));

        // ---- sub-panel columns into the register window: a[r][k] = column j + k at step j (sliding) ----
        T a[RPT][SW];
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const T* prow = s_panel + (tid + r * TPB) * LD + sb;
            const bool dead = kpr[r] == PP_NOPOSROW;
#pragma unroll
            for (int c = 0; c < SW / VEC; ++c) {
                const V16 v = *reinterpret_cast<const V16*>(prow + c * VEC);
#pragma unroll
                for (int e = 0; e < VEC; ++e) a[r][c * VEC + e] = dead ? T(0) : v.v[e];  // retired rows: key 0 forever
            }
        }

        PP_COARSE(1);

        // the rest of the pivot row of column `pend_c` (its columns sb.. : multipliers, U, parked part) to every CTA's s_piv;
        // runs on the warps that idle while warp 0 exchanges records, one block barrier after the row was made current
        auto push_rest = [&]() {
            if (lane < nrest) {
                const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(s_panel + pend_lrow * LD + sb + lane * VEC);
                const unsigned dst = pp_smem_u32(s_piv + pend_c * W + sb + lane * VEC);
                const unsigned pbar = pp_smem_u32(&s_pbar[sp]);
                for (int peer = warp - 1; peer < C; peer += NW - 1) pp_push16(pp_mapa(dst, (unsigned)peer), v, pp_mapa(pbar, (unsigned)peer));
            }
        };

        int c = 0;
#pragma unroll 1
        for (; c < SW; ++c) {
            const int j = sb + c;
            if (j >= w) break;  // uniform
            const int parity = j & 1;
            const int left = SW - c;  // window slots still inside the sub-panel
            if (timing == 1 && rank == 0 && tid == 0) tprev = clock64();

            // (1) thread candidate over its rows (retired rows hold 0 / PP_NOPOSROW), warp candidate
            KT bkey = K::of(a[0][0]);
            unsigned bpr = kpr[0];
            int br = 0;
#pragma unroll
            for (int r = 1; r < RPT; ++r) {
                const KT kr = K::of(a[r][0]);
                const bool better = kr > bkey || (kr == bkey && kpr[r] < bpr);
                bkey = better ? kr : bkey;
                bpr = better ? kpr[r] : bpr;
                br = better ? r : br;
            }
            const uint32_t bhi = (bpr != PP_NOPOSROW) ? 1u + ((sizeof(KT) == 8) ? (uint32_t)((unsigned long long)bkey >> 32) : (uint32_t)bkey) : 0u;
            auto write_record = [&]() {
                unsigned char* rec = s_wc + warp * REC;
                *reinterpret_cast<unsigned long long*>(rec + 8) = (unsigned long long)bhi | ((unsigned long long)bpr << 32);
#pragma unroll
                for (int r = 0; r < RPT; ++r) {
                    if (br == r) {
#pragma unroll
                        for (int q = 0; q < SW / VEC; ++q) {
                            V16 v;
#pragma unroll
                            for (int e = 0; e < VEC; ++e) v.v[e] = a[r][q * VEC + e];
                            *reinterpret_cast<V16*>(rec + 16 + q * 16) = v;
                        }
                    }
                }
            };
            bool ex1;
            const int wsrc = pp_argmax_fast(bhi, ex1);
            if (lane == wsrc) write_record();
            if (ex1) {  // rare: coarse tie inside the warp (or no live row): the exact choice overwrites the record
                const int w2 = pp_argmax_hdr<KT>(bhi, bpr, [&]() { return bkey; });
                if (w2 != wsrc && lane == w2) write_record();
            }
            PP_STAMP(0);
            __syncthreads();
            PP_STAMP(1);
            if (timing == 3 && rank == 0 && tid == 0) tprev = clock64();

            if (warp == 0) {
                // (2) warp 0: CTA candidate among the NW warp records (fast arg-max, exact only on a coarse tie), then the record to every CTA:
                //     window chunks at once, the header chunk when warp 1 has delivered the reciprocals
                unsigned long long hy = (unsigned long long)PP_NOPOSROW << 32;
                if (lane < NW) hy = *reinterpret_cast<const unsigned long long*>(s_wc + lane * REC + 8);
                bool ex2;
                int cw = pp_argmax_fast((uint32_t)hy, ex2);
                if (cw >= NW) cw = 0;
                const unsigned bar = pp_smem_u32(&s_mbar[parity]);
                // only the window chunks that still hold live columns travel (slots >= left are never read by anybody)
                const int nch = 1 + (left + VEC - 1) / VEC;
                if (lane == 0) pp_expect_tx(bar, (unsigned)(C * nch * 16));
                const unsigned dst = push_dst0 + (unsigned)(parity * PP_MAXC * REC);
                const unsigned rbar = push_bar0 + (unsigned)(parity * 8);
                constexpr int NQ = (NCH + 1) / 2;
                ulonglong2 wv[NQ];  // this lane's window chunks of the (speculative) winner, loaded under the tie check
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    const int ch = my_ch0 + 2 * q;
                    wv[q] = *reinterpret_cast<const ulonglong2*>(s_wc + cw * REC + (ch < NCH ? ch : 0) * 16);
                }
                if (ex2) {  // rare: coarse tie among the warp records
                    cw = pp_argmax_hdr<KT>((uint32_t)hy, (unsigned)(hy >> 32),
                                           [&]() { return lane < NW ? K::of(*reinterpret_cast<const T*>(s_wc + lane * REC + 16)) : (KT)0; });
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        const int ch = my_ch0 + 2 * q;
                        wv[q] = *reinterpret_cast<const ulonglong2*>(s_wc + cw * REC + (ch < NCH ? ch : 0) * 16);
                    }
                }
                PP_FINE(0);
                PP_FINE(1);
                const unsigned char* src = s_wc + cw * REC;
                if (my_peer < C) {
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        const int ch = my_ch0 + 2 * q;
                        if (ch > 0 && ch < nch) pp_push16(dst + ch * 16, wv[q], rbar);
                    }
                }
                PP_FINE(2);
                asm volatile("bar.sync 1, 64;" ::: "memory");  // warp 1 has written 1 / candidate into every warp record
                PP_FINE(3);
                if (my_peer < C && my_ch0 == 0) pp_push16(dst, *reinterpret_cast<const ulonglong2*>(src), rbar);
                PP_FINE(4);
            } else {
                if (warp == 1) {
                    // 1 / candidate of every warp record (A::one() / pivot, getrf.rs:76), under warp 0's reduction
                    if (lane < NW) {
                        const T w0 = *reinterpret_cast<const T*>(s_wc + lane * REC + 16);
                        *reinterpret_cast<unsigned long long*>(s_wc + lane * REC) = RC::bits(T(1) / w0);
                    }
                    asm volatile("bar.arrive 1, 64;" ::: "memory");
                }
                if (pend_c >= 0) push_rest();
            }
            pend_c = -1;
            PP_STAMP(2);

            // (3) every warp: wait for the C records, decide the winner from the headers, read its header and window
            pp_mbar_wait(pp_smem_u32(&s_mbar[parity]), (unsigned)(j >> 1) & 1u);
            PP_STAMP(3);
            PP_FINE(5);
            if (timing == 3 && rank == 0 && tid == 0) tcols += 1;
            const unsigned char* cbase = s_cand + parity * PP_MAXC * REC;
            unsigned long long hy4 = (unsigned long long)PP_NOPOSROW << 32;
            if (lane < C) hy4 = *reinterpret_cast<const unsigned long long*>(cbase + lane * REC + 8);
            bool ex4;
            int gw = pp_argmax_fast((uint32_t)hy4, ex4);
            if (gw >= C) gw = 0;
            const unsigned char* wrec = cbase + gw * REC;
            ulonglong2 wh = *reinterpret_cast<const ulonglong2*>(wrec);  // (speculative: loaded under the tie check)
            T u[SW];  // the pivot row's window: u[k] = its entry of column j + k (garbage beyond the sub-panel, never used)
#pragma unroll
            for (int q = 0; q < SW / VEC; ++q) {
                const V16 v = *reinterpret_cast<const V16*>(wrec + 16 + q * 16);
#pragma unroll
                for (int e = 0; e < VEC; ++e) u[q * VEC + e] = v.v[e];
            }
            if (ex4) {  // rare: coarse tie among the CTA records
                gw = pp_argmax_hdr<KT>((uint32_t)hy4, (unsigned)(hy4 >> 32),
                                       [&]() { return lane < C ? K::of(*reinterpret_cast<const T*>(cbase + lane * REC + 16)) : (KT)0; });
                if (gw >= C) gw = 0;
                wrec = cbase + gw * REC;
                wh = *reinterpret_cast<const ulonglong2*>(wrec);
#pragma unroll
                for (int q = 0; q < SW / VEC; ++q) {
                    const V16 v = *reinterpret_cast<const V16*>(wrec + 16 + q * 16);
#pragma unroll
                    for (int e = 0; e < VEC; ++e) u[q * VEC + e] = v.v[e];
                }
            }
            const unsigned gpr = (unsigned)(wh.y >> 32);
            const int gpos = (int)(gpr >> 12);
            const bool sing = (K::of(u[0]) == (KT)0);  // nothing exceeded 0 (iamax.rs:10-19): zero (or NaN) pivot
            const T recip = RC::from(wh.x);
            if (rank == 0 && tid == 0) {
                ipiv[j] = row_base + gpos;
                if (sing) *info = step_base + j;  // last zero-pivot step wins (getrf.rs:72-73)
            }
            if (gw == rank && npark > 0) {
                pend_c = c;
                pend_lrow = (int)(gpr & 0xfffu);
            }
            PP_STAMP(4);

            // (4) rank-1 update of the live rows fused with the window shift; the winner retires
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                T* prow = s_panel + (tid + r * TPB) * LD + j;
                const bool won = (kpr[r] == gpr);  // positions are unique among live rows
                if (pos[r] == j) {  // the row at position j trades places with the winner
                    pos[r] = gpos;
                    kpr[r] = (kpr[r] & 0xfffu) | ((unsigned)gpos << 12);
                }
                if (won) {  // (also when it already sat at position j)
                    pos[r] = j;
                    kpr[r] = PP_NOPOSROW;
#pragma unroll
                    for (int k = 0; k < SW; ++k)
                        if (k < left) prow[k] = a[r][k];  // its U entries to the panel (its multipliers are already there)
                    a[r][0] = T(0);
                } else if (kpr[r] != PP_NOPOSROW) {
                    if (!sing) {
                        const T l = a[r][0] * recip;
                        prow[0] = l;
#pragma unroll
                        for (int k = 1; k < SW; ++k) a[r][k - 1] = pp_fnma(l, u[k], a[r][k]);
                    } else {
                        prow[0] = a[r][0];
#pragma unroll
                        for (int k = 1; k < SW; ++k) a[r][k - 1] = a[r][k];
                    }
                    a[r][SW - 1] = T(0);
                }
            }
            PP_STAMP(5);
            if (timing == 1 && rank == 0 && tid == 0) tcols += 1;
        }
        PP_COARSE(2);

        if (npark > 0) {
            __syncthreads();  // the last pivot row is current in its owner's panel
            if (warp != 0 && pend_c >= 0) push_rest();
            pend_c = -1;
            pp_mbar_wait(pp_smem_u32(&s_pbar[sp]), (unsigned)(sb >> 4) & 1u);  // the SW pivot rows (columns sb..) have landed
            PP_COARSE(3);
            // U12 = L11^-1 * P12 by forward substitution, one thread per parked column (tiny, redundantly per CTA)
            if (tid < npark) {
                T l11[SW * (SW - 1) / 2];
                T uc[SW];
#pragma unroll
                for (int i = 1; i < SW; ++i) {
#pragma unroll
                    for (int k = 0; k < i; ++k) l11[i * (i - 1) / 2 + k] = s_piv[i * W + sb + k];
                }
#pragma unroll
                for (int i = 0; i < SW; ++i) uc[i] = s_piv[i * W + c1 + tid];
#pragma unroll
                for (int i = 1; i < SW; ++i) {
#pragma unroll
                    for (int k = 0; k < i; ++k) uc[i] -= l11[i * (i - 1) / 2 + k] * uc[k];
                }
#pragma unroll
                for (int i = 0; i < SW; ++i) s_piv[i * W + c1 + tid] = uc[i];
            }
            __syncthreads();
            PP_COARSE(4);
            // parked columns: pivot rows of this sub-panel take their U12 row; live rows get the rank-8 update
            T lm[RPT][SW];
            bool livr[RPT];
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                T* prow = s_panel + (tid + r * TPB) * LD;
                const int p = pos[r];
                livr[r] = p >= sb + SW;
                if (p >= sb && p < sb + SW) {
                    const T* urow12 = s_piv + (p - sb) * W;
                    for (int k = c1; k < W; k += VEC) *reinterpret_cast<V16*>(prow + k) = *reinterpret_cast<const V16*>(urow12 + k);
                }
#pragma unroll
                for (int cc = 0; cc < SW / VEC; ++cc) {
                    const V16 v = *reinterpret_cast<const V16*>(prow + sb + cc * VEC);
#pragma unroll
                    for (int e = 0; e < VEC; ++e) lm[r][cc * VEC + e] = v.v[e];
                }
            }
#pragma unroll 2
            for (int k = c1; k < W; k += VEC) {
                V16 uu[SW];
#pragma unroll
                for (int i = 0; i < SW; ++i) uu[i] = *reinterpret_cast<const V16*>(s_piv + i * W + k);
#pragma unroll
                for (int r = 0; r < RPT; ++r) {
                    if (livr[r]) {
                        T* prow = s_panel + (tid + r * TPB) * LD;
                        V16 x = *reinterpret_cast<const V16*>(prow + k);
#pragma unroll
                        for (int i = 0; i < SW; ++i) {
#pragma unroll
                            for (int e = 0; e < VEC; ++e) x.v[e] -= lm[r][i] * uu[i].v[e];
                        }
                        *reinterpret_cast<V16*>(prow + k) = x;
                    }
                }
            }
            __syncthreads();
            PP_COARSE(5);
            if (timing == 2 && rank == 0 && tid == 0) tcols += 1;
        }
    }

    // ---- rows to their final positions ----
    if (vec_ok) {
#pragma unroll
        for (int r = 0; r < RPT; ++r) s_pos[tid + r * TPB] = pos[r];
        __syncthreads();
        constexpr int CPR = W / VEC;
        for (int c = tid; c < ROWS * CPR; c += TPB) {
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
        g_pp_timing[0] += tacc0;
        g_pp_timing[1] += tacc1;
        g_pp_timing[2] += tacc2;
        g_pp_timing[3] += tacc3;
        g_pp_timing[4] += tacc4;
        g_pp_timing[5] += tacc5;
        g_pp_timing[6] += tcols;
        g_pp_timing[7] += clock64() - tstart;
    }
#undef PP_STAMP
#undef PP_COARSE
#undef PP_FINE
}

template <class T, int RPT, int W, int ROWS>
int launch_push(int64_t rows, int64_t w, T* d_a, int64_t lda, int32_t* d_ipiv, int32_t row_base, int32_t* d_info, int32_t step_base,
                cudaStream_t s) {
    auto kern = panel_push_kernel<T, RPT, W, ROWS>;
    constexpr int TPB = ROWS / RPT;
    const size_t smem = PPSmem<T, W, ROWS>::total;
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
    const int need = (int)((rows + ROWS - 1) / ROWS);
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
    LAIR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, d_a, (long long)lda, (int)rows, (int)w, d_ipiv, (int)row_base, d_info, (int)step_base,
                                       (int)ctx().opt.panel_timing));
    LAIR_LAUNCH_CHECK();
    return LAIR_B200_OK;
}

// rows per CTA: 32-wide panels 512 (f32 also 1024: 16 384 rows in one cluster); 64-wide panels 256 (f64) / 512 (f32)
template <class T> constexpr int pp_rows64() { return sizeof(T) == 8 ? 256 : 512; }
constexpr int PP_ROWS32 = 512;

}  // namespace

// Widest panel one launch takes for `rows` rows: 64 (option panel_w64) while the rows fit 16 CTAs of the 64-wide
// layout, else 32, else 0.
template <class T>
int panel_push_max_width(int64_t rows) {
    if (ctx().opt.panel_w64 != 0 && rows <= (int64_t)PP_MAXC * pp_rows64<T>()) return 64;
    if (rows <= (int64_t)PP_MAXC * PP_ROWS32) return 32;
    if (sizeof(T) == 4 && rows <= (int64_t)PP_MAXC * 1024) return 32;
    return 0;
}

// Returns LAIR_B200_ERR_UNSUPPORTED (without setting an error) when the panel does not fit one cluster.
template <class T>
int panel_push_dev(int64_t rows, int64_t w, T* d_a, int64_t lda, int32_t* d_ipiv, int32_t row_base, int32_t* d_info, int32_t step_base,
                   cudaStream_t s) {
    // rows per thread: option panel_rpt forces 1 / 2 / 4; 0 (default) picks the measured best (profiles/r2p_probe_panel_push.jsonl):
    // one row per thread for 64-wide panels and for f32, two for 32-wide f64 panels
    int rpt = (int)ctx().opt.panel_rpt;
    if (w > 32) {
        if (w > 64 || ctx().opt.panel_w64 == 0 || rows > (int64_t)PP_MAXC * pp_rows64<T>()) return LAIR_B200_ERR_UNSUPPORTED;
        if (rpt == 0) rpt = 1;
        if (rpt == 1) return launch_push<T, 1, 64, pp_rows64<T>()>(rows, w, d_a, lda, d_ipiv, row_base, d_info, step_base, s);
        if constexpr (sizeof(T) == 4) {
            if (rpt == 4) return launch_push<T, 4, 64, pp_rows64<T>()>(rows, w, d_a, lda, d_ipiv, row_base, d_info, step_base, s);
        }
        return launch_push<T, 2, 64, pp_rows64<T>()>(rows, w, d_a, lda, d_ipiv, row_base, d_info, step_base, s);
    }
    if (rows <= (int64_t)PP_MAXC * PP_ROWS32) {
        if (rpt == 0) rpt = sizeof(T) == 4 ? 1 : 2;
        if (rpt == 1) return launch_push<T, 1, 32, PP_ROWS32>(rows, w, d_a, lda, d_ipiv, row_base, d_info, step_base, s);
        if (rpt == 4) return launch_push<T, 4, 32, PP_ROWS32>(rows, w, d_a, lda, d_ipiv, row_base, d_info, step_base, s);
        return launch_push<T, 2, 32, PP_ROWS32>(rows, w, d_a, lda, d_ipiv, row_base, d_info, step_base, s);
    }
    if constexpr (sizeof(T) == 4) {
        if (rows <= (int64_t)PP_MAXC * 1024) {
            if (rpt == 4) return launch_push<T, 4, 32, 1024>(rows, w, d_a, lda, d_ipiv, row_base, d_info, step_base, s);
            return launch_push<T, 2, 32, 1024>(rows, w, d_a, lda, d_ipiv, row_base, d_info, step_base, s);
        }
    }
    return LAIR_B200_ERR_UNSUPPORTED;
}

int panel_push_timing(long long* out8, bool clear) {
    LAIR_CUDA_CHECK(cudaDeviceSynchronize());
    LAIR_CUDA_CHECK(cudaMemcpyFromSymbol(out8, g_pp_timing, 8 * sizeof(long long)));
    if (clear) {
        long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        LAIR_CUDA_CHECK(cudaMemcpyToSymbol(g_pp_timing, z, sizeof(z)));
    }
    return LAIR_B200_OK;
}

template int panel_push_dev<float>(int64_t, int64_t, float*, int64_t, int32_t*, int32_t, int32_t*, int32_t, cudaStream_t);
template int panel_push_dev<double>(int64_t, int64_t, double*, int64_t, int32_t*, int32_t, int32_t*, int32_t, cudaStream_t);
template int panel_push_max_width<float>(int64_t);
template int panel_push_max_width<double>(int64_t);

}  // namespace lair
