// Panel factorization: LU with partial pivoting of a tall-skinny (rows x w, w <= 32) block,
// the latency-critical step of the blocked getrf (one dependent arg-max per column).
//
// Design (B200-first, not a translation of the reference's column loop):
//   * The whole panel lives ON CHIP for the duration of the kernel: every thread owns RPT
//     matrix rows of W columns in REGISTERS; the grid is sized so all CTAs are co-resident.
//   * Row interchanges are logical: a row never moves, only its position `pos` changes; rows
//     are written to their final positions once, at the end.
//   * Per column: thread-local candidate -> warp arg-max with REDUX (first maximum =
//     lowest position among equal |x|, NaN never wins; src/blas/iamax.rs:6-21) -> CTA
//     candidate -> ONE grid-wide exchange through global memory using a flag-in-data
//     ("LL") protocol: every 8-byte payload travels with a sequence number in the same
//     16-byte store, so there is no separate flag, no fence and no grid barrier.  Each CTA
//     publishes its candidate ROW speculatively; every CTA then reads all candidates,
//     picks the same winner and reads the winner's row as the U row of the rank-1 update.
//   * The update a[k] -= l * u[k] runs from registers with the multiplier formed by
//     reciprocal-multiply like the reference (src/lapack/getrf.rs:76-81).
// Roofline: HBM traffic is 2 * rows * w * sizeof(T) per launch, but the kernel is bound by
// the per-column exchange latency (two L2 round trips), not bandwidth.
#include <climits>

#include "common.cuh"
#include "pivot_key.cuh"

namespace lair {
namespace {

constexpr int PANEL_TPB = 256;
constexpr int PANEL_GMAX = 1024;       // max CTAs of one panel launch
constexpr int PANEL_SLOT_CHUNKS = 40;  // 16-byte chunks per (parity, cta) slot (33 used)
constexpr size_t PANEL_WS_BYTES = (size_t)2 * PANEL_GMAX * PANEL_SLOT_CHUNKS * 16 + 256;
constexpr int PANEL_SPIN_LIMIT = 4000000;  // ~1 s of polling before giving up (never hang the GPU)

__device__ __forceinline__ void st_ll(uint4* p, unsigned long long payload, uint32_t seq) {
    asm volatile("st.relaxed.gpu.global.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "r"((uint32_t)payload), "r"(seq),
                 "r"((uint32_t)(payload >> 32)), "r"(seq)
                 : "memory");
}
__device__ __forceinline__ bool ld_ll_try(const uint4* p, uint32_t seq, unsigned long long& payload) {
    uint32_t a, b, c, d;
    asm volatile("ld.relaxed.gpu.global.v4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
    payload = ((unsigned long long)c << 32) | a;
    return b == seq && d == seq;
}
__device__ __forceinline__ unsigned long long ld_ll_wait(const uint4* p, uint32_t seq, int* err) {
    unsigned long long v = 0;
    int spins = 0;
    while (!ld_ll_try(p, seq, v)) {
        if (++spins > PANEL_SPIN_LIMIT) {
            atomicExch(err, 1);
            break;
        }
    }
    return v;
}
__device__ __forceinline__ void ld_ll_wait2(const uint4* p0, const uint4* p1, uint32_t seq, unsigned long long& v0,
                                            unsigned long long& v1, int* err) {
    int spins = 0;
    for (;;) {
        bool ok0 = ld_ll_try(p0, seq, v0);
        bool ok1 = ld_ll_try(p1, seq, v1);
        if (ok0 && ok1) break;
        if (++spins > PANEL_SPIN_LIMIT) {
            atomicExch(err, 1);
            break;
        }
    }
}

template <class T> __device__ __forceinline__ T payload_elem(unsigned long long pv, int sub);
template <> __device__ __forceinline__ double payload_elem<double>(unsigned long long pv, int) { return __longlong_as_double((long long)pv); }
template <> __device__ __forceinline__ float payload_elem<float>(unsigned long long pv, int sub) {
    return __uint_as_float(sub ? (uint32_t)(pv >> 32) : (uint32_t)pv);
}

template <class T, int W, int RPT, int MINB>
__global__ void __launch_bounds__(PANEL_TPB, MINB)
panel_kernel(T* __restrict__ A, long long lda, int M, int w, int32_t* __restrict__ ipiv, int row_base,
             int32_t* __restrict__ info, int step_base, uint4* __restrict__ ws, uint32_t seq_base, int* __restrict__ err) {
    using K = PivotKey<T>;
    using KT = typename K::type;
    constexpr int EPC = 8 / sizeof(T);  // elements per 8-byte payload
    constexpr int NCH = W / EPC;        // payload chunks per row
    static_assert(NCH <= 32 && NCH + 1 <= PANEL_SLOT_CHUNKS, "row must fit one warp-wide LL store");
    constexpr int NW = PANEL_TPB / 32;
    constexpr int VEC = 16 / sizeof(T);
    struct alignas(16) V16 { T v[VEC]; };

    __shared__ KT s_wkey[NW];
    __shared__ int s_wpos[NW];
    __shared__ __align__(16) T s_wrow[NW][W];
    __shared__ __align__(16) T s_urow[W];
    __shared__ T s_recip;
    __shared__ int s_gpos, s_sing;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = blockIdx.x, G = gridDim.x;

    // ---- load: RPT rows per thread, interleaved over the grid ----
    T a[RPT][W];
    int pos[RPT];
    const bool vec_ok = (w == W) && ((lda % VEC) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int row = (r * G + cta) * PANEL_TPB + tid;
        pos[r] = row < M ? row : -1;
        if (row < M) {
            const T* p = A + (long long)row * lda;
            if (vec_ok) {
#pragma unroll
                for (int c = 0; c < W / VEC; ++c) {
                    V16 v = *reinterpret_cast<const V16*>(p + c * VEC);
#pragma unroll
                    for (int e = 0; e < VEC; ++e) a[r][c * VEC + e] = v.v[e];
                }
            } else {
#pragma unroll
                for (int c = 0; c < W; ++c) a[r][c] = (c < w) ? p[c] : T(0);
            }
        } else {
#pragma unroll
            for (int c = 0; c < W; ++c) a[r][c] = T(0);
        }
    }

#pragma unroll
    for (int j = 0; j < W; ++j) {
        if (j >= w) break;
        const uint32_t seq = seq_base + (uint32_t)j + 1u;
        const int parity = j & 1;
        const int cj = j / EPC;

        // (1) thread candidate over its live rows, then warp candidate
        KT bkey = 0;
        int bpos = INT_MAX, br = 0;
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const bool live = pos[r] >= j;
            const KT key = live ? K::of(a[r][j]) : (KT)0;
            const int p = live ? pos[r] : INT_MAX;
            if (key > bkey || (key == bkey && p < bpos)) {
                bkey = key;
                bpos = p;
                br = r;
            }
        }
        const KT wmax = K::warp_max(bkey);
        const bool cand = (bkey == wmax);
        const unsigned wpos = __reduce_min_sync(kFullMask, cand ? (unsigned)bpos : 0xffffffffu);
        const bool own = cand && ((unsigned)bpos == wpos) && (wpos != (unsigned)INT_MAX);
        if (own) {
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                if (br == r) {
#pragma unroll
                    for (int c = 0; c < W / VEC; ++c) {
                        V16 v;
#pragma unroll
                        for (int e = 0; e < VEC; ++e) v.v[e] = a[r][c * VEC + e];
                        *reinterpret_cast<V16*>(&s_wrow[warp][c * VEC]) = v;
                    }
                }
            }
        }
        if (lane == 0) {
            s_wkey[warp] = wmax;
            s_wpos[warp] = (int)wpos;
        }
        __syncthreads();

        // (2)+(3) warp 0: CTA candidate -> publish -> read every CTA's candidate -> winner row
        if (warp == 0) {
            const KT k8 = lane < NW ? s_wkey[lane] : (KT)0;
            const unsigned p8 = lane < NW ? (unsigned)s_wpos[lane] : 0xffffffffu;
            const KT cmax = K::warp_max(k8);
            const bool c8 = (lane < NW) && (k8 == cmax);
            const unsigned cpos = __reduce_min_sync(kFullMask, c8 ? p8 : 0xffffffffu);
            const int bw = __ffs(__ballot_sync(kFullMask, c8 && p8 == cpos)) - 1;
            uint4* myslot = ws + (size_t)(parity * PANEL_GMAX + cta) * PANEL_SLOT_CHUNKS;
            if (lane < NCH) st_ll(myslot + lane, reinterpret_cast<const unsigned long long*>(&s_wrow[bw][0])[lane], seq);
            if (lane == 0) st_ll(myslot + NCH, (unsigned long long)cpos, seq);

            KT gk = 0;
            unsigned gp = 0xffffffffu;
            int gidx = 0;
            T gval = T(0);
            for (int g = lane; g < G; g += 32) {
                const uint4* sl = ws + (size_t)(parity * PANEL_GMAX + g) * PANEL_SLOT_CHUNKS;
                unsigned long long pv, pp;
                ld_ll_wait2(sl + cj, sl + NCH, seq, pv, pp, err);
                const T val = payload_elem<T>(pv, j % EPC);
                const KT key = K::of(val);
                const unsigned p = (unsigned)pp;
                if (key > gk || (key == gk && p < gp)) {
                    gk = key;
                    gp = p;
                    gidx = g;
                    gval = val;
                }
            }
            const KT gmax = K::warp_max(gk);
            const bool c2 = (gk == gmax);
            const unsigned gpos = __reduce_min_sync(kFullMask, c2 ? gp : 0xffffffffu);
            const int wl = __ffs(__ballot_sync(kFullMask, c2 && gp == gpos)) - 1;
            const int gw = __shfl_sync(kFullMask, gidx, wl);
            const T pivval = __shfl_sync(kFullMask, gval, wl);
            const uint4* wsl = ws + (size_t)(parity * PANEL_GMAX + gw) * PANEL_SLOT_CHUNKS;
            if (lane >= cj && lane < NCH)
                reinterpret_cast<unsigned long long*>(s_urow)[lane] = ld_ll_wait(wsl + lane, seq, err);
            if (lane == 0) {
                s_gpos = (int)gpos;
                s_sing = (gmax == 0);
                s_recip = (gmax == 0) ? T(0) : T(1) / pivval;  // A::one() / pivot (getrf.rs:76)
                if (cta == 0) {
                    ipiv[j] = row_base + (int)gpos;
                    if (gmax == 0) *info = step_base + j;  // last zero-pivot step wins (getrf.rs:72-73)
                }
            }
        }
        __syncthreads();

        // (4) logical swap and rank-1 update from registers
        const int gpos = s_gpos;
        const bool sing = s_sing != 0;
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const bool was_j = (pos[r] == j), was_w = (pos[r] == gpos);
            if (was_j) pos[r] = gpos;
            if (was_w) pos[r] = j;
        }
        if (!sing) {
            const T recip = s_recip;
            T u[W];
#pragma unroll
            for (int c = (j + 1) / VEC; c < W / VEC; ++c) {
                V16 v = *reinterpret_cast<const V16*>(&s_urow[c * VEC]);
#pragma unroll
                for (int e = 0; e < VEC; ++e) u[c * VEC + e] = v.v[e];
            }
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                if (pos[r] > j) {
                    const T l = a[r][j] * recip;
                    a[r][j] = l;
#pragma unroll
                    for (int k = j + 1; k < W; ++k) a[r][k] -= l * u[k];
                }
            }
        }
    }

    // ---- rows to their final positions ----
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        if (pos[r] >= 0) {
            T* p = A + (long long)pos[r] * lda;
            if (vec_ok) {
#pragma unroll
                for (int c = 0; c < W / VEC; ++c) {
                    V16 v;
#pragma unroll
                    for (int e = 0; e < VEC; ++e) v.v[e] = a[r][c * VEC + e];
                    *reinterpret_cast<V16*>(p + c * VEC) = v;
                }
            } else {
#pragma unroll
                for (int c = 0; c < W; ++c)
                    if (c < w) p[c] = a[r][c];
            }
        }
    }
}

template <class T, int W, int RPT, int MINB>
struct PanelVariant {
    static int capacity_rows(int* grid_cap) {
        static int cached_blocks = -1;
        static uint64_t seen_epoch = 0;
        if (stale_for_context(seen_epoch)) cached_blocks = -1;
        if (cached_blocks < 0) {
            int b = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, panel_kernel<T, W, RPT, MINB>, PANEL_TPB, 0) != cudaSuccess) b = 0;
            cached_blocks = b;
        }
        long long g = (long long)cached_blocks * ctx().sm_count;
        if (g > PANEL_GMAX) g = PANEL_GMAX;
        if (grid_cap) *grid_cap = (int)g;
        long long rows = g * PANEL_TPB * RPT;
        return rows > INT_MAX ? INT_MAX : (int)rows;
    }
    static int launch(int64_t rows, int64_t w, T* d_a, int64_t lda, int32_t* d_ipiv, int32_t row_base, int32_t* d_info,
                      int32_t step_base, cudaStream_t s) {
        Context& c = ctx();
        if (!c.panel_ws) {
            LAIR_CUDA_CHECK(cudaMalloc(&c.panel_ws, PANEL_WS_BYTES));
            LAIR_CUDA_CHECK(cudaMemset(c.panel_ws, 0, PANEL_WS_BYTES));
            c.panel_ws_bytes = PANEL_WS_BYTES;
            c.panel_seq = 0;
        }
        if (c.panel_seq > 0xF0000000u) {  // sequence numbers about to wrap: start over on clean slots
            LAIR_CUDA_CHECK(cudaDeviceSynchronize());
            LAIR_CUDA_CHECK(cudaMemset(c.panel_ws, 0, PANEL_WS_BYTES));
            c.panel_seq = 0;
        }
        const int per_cta = PANEL_TPB * RPT;
        const int grid = (int)((rows + per_cta - 1) / per_cta);
        uint4* ws = reinterpret_cast<uint4*>(c.panel_ws);
        int* err = c.d_fault;  // an exchange that times out raises the context's fault word (check_fault)
        ProfScope prof(kProfPanel, s, 2.0 * (double)rows * (double)w * sizeof(T));
        panel_kernel<T, W, RPT, MINB><<<grid, PANEL_TPB, 0, s>>>(d_a, (long long)lda, (int)rows, (int)w, d_ipiv, row_base, d_info,
                                                                 step_base, ws, c.panel_seq, err);
        c.panel_seq += 64;
        LAIR_LAUNCH_CHECK();
        return LAIR_B200_OK;
    }
};

// Variant tables: widest panel first.
template <class T> struct PanelTable;
template <> struct PanelTable<double> {
    using V0 = PanelVariant<double, 32, 1, 2>;
    using V1 = PanelVariant<double, 8, 4, 2>;
    static constexpr int W0 = 32, W1 = 8;
};
template <> struct PanelTable<float> {
    using V0 = PanelVariant<float, 32, 1, 2>;
    using V1 = PanelVariant<float, 16, 4, 2>;
    static constexpr int W0 = 32, W1 = 16;
};

}  // namespace

template <class T>
int panel_max_width(int64_t rows) {
    using PT = PanelTable<T>;
    if (ctx().opt.panel_cluster >= 3) {
        const int wb = panel_push_max_width<T>(rows);
        if (wb > 0) return wb;
    }
    if (ctx().opt.panel_cluster >= 2 && rows <= panel_cluster_max_rows()) {
        const int wb = panel_blocked_max_width<T>(rows);
        if (wb > 0) return wb;
    }
    if (ctx().opt.panel_cluster && rows <= panel_cluster_max_rows()) return 32;
    if (rows <= PT::V0::capacity_rows(nullptr)) return PT::W0;
    if (rows <= PT::V1::capacity_rows(nullptr)) return PT::W1;
    return 0;
}

template <class T>
int panel_dev(int64_t rows, int64_t w, T* d_a, int64_t lda, int32_t* d_ipiv, int32_t row_base, int32_t* d_info,
              int32_t step_base, cudaStream_t s) {
    using PT = PanelTable<T>;
    LAIR_REQUIRE(rows >= 1 && w >= 1 && w <= rows, "panel: bad shape rows=%lld w=%lld", (long long)rows, (long long)w);
    if (ctx().opt.panel_cluster >= 3) {  // fourth generation: records carry the register window, every warp decides
        int st = panel_push_dev<T>(rows, w, d_a, lda, d_ipiv, row_base, d_info, step_base, s);
        if (st != LAIR_B200_ERR_UNSUPPORTED) return st;
    }
    if (ctx().opt.panel_cluster && rows <= panel_cluster_max_rows() && (w <= 32 || ctx().opt.panel_cluster >= 2)) {
        // 2 (default): in-kernel blocked cluster kernel; 1: one-row-per-thread cluster kernel
        int st = (ctx().opt.panel_cluster >= 2) ? panel_blocked_dev<T>(rows, w, d_a, lda, d_ipiv, row_base, d_info, step_base, s)
                                                : panel_cluster_dev<T>(rows, w, d_a, lda, d_ipiv, row_base, d_info, step_base, s);
        if (st != LAIR_B200_ERR_UNSUPPORTED) return st;
    }
    if (w <= PT::W0 && rows <= PT::V0::capacity_rows(nullptr))
        return PT::V0::launch(rows, w, d_a, lda, d_ipiv, row_base, d_info, step_base, s);
    if (w <= PT::W1 && rows <= PT::V1::capacity_rows(nullptr))
        return PT::V1::launch(rows, w, d_a, lda, d_ipiv, row_base, d_info, step_base, s);
    set_error("panel: %lld rows x %lld columns exceeds the on-chip panel capacity of this device", (long long)rows,
              (long long)w);
    return LAIR_B200_ERR_UNSUPPORTED;
}


template int panel_max_width<float>(int64_t);
template int panel_max_width<double>(int64_t);
template int panel_dev<float>(int64_t, int64_t, float*, int64_t, int32_t*, int32_t, int32_t*, int32_t, cudaStream_t);
template int panel_dev<double>(int64_t, int64_t, double*, int64_t, int32_t*, int32_t, int32_t*, int32_t, cudaStream_t);

}  // namespace lair
