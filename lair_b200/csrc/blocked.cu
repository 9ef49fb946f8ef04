// Blocked LU with partial pivoting on a device-resident row-major matrix, and the blocked
// multi-RHS solve.  Shape of the algorithm = the reference's recursive variant
// (src/lapack/getrf.rs:216-322: factor left / laswp / trsm / gemm / factor right / laswp),
// which the reference's own test pins to the same pivots and L\U as its live unblocked loops
// (getrf.rs:381-403); here the recursion bottoms out in the on-chip panel kernel (panel.cu)
// instead of a single column, and the outer level is an iterative right-looking sweep of
// width nb so the trailing update is one large DMMA GEMM per block column.
#include "common.cuh"

namespace lair {
namespace {

__global__ void set_i32_kernel(int32_t* p, int32_t v) { *p = v; }

template <class T>
struct Factor {
    T* A;
    int64_t lda, m, n;
    int32_t* ipiv;
    int32_t* info;
    cudaStream_t s;  // the caller's stream: everything is ordered after / visible on it
    const ColumnFeed* feed = nullptr;  // columns still arriving from the host (run() only)
    RowDrain* drain = nullptr;         // finished rows leave for the host while the sweep runs (run() only)

    T* at(int64_t r, int64_t c) const { return A + r * lda + c; }

    // columns [c0, c1) of the whole matrix get the interchanges ipiv[k0..k1)
    int swap_cols(int64_t c0, int64_t c1, int64_t k0, int64_t k1, cudaStream_t st) const {
        if (c1 <= c0 || k1 <= k0) return LAIR_B200_OK;
        return laswp_dev<T>(c1 - c0, A + c0, lda, k0, k1, ipiv, st);
    }

    // laswp(ipiv[k0..k0+k)) on columns [c0, c1), then the unit-lower solve of their rows k0..k0+k
    // against the k x k block at (k0, lc0): one fused launch for narrow blocks, else laswp + trsm.
    int swap_solve(int64_t k0, int64_t k, int64_t lc0, int64_t c0, int64_t c1, cudaStream_t st) const {
        if (c1 <= c0 || k <= 0) return LAIR_B200_OK;
        const int64_t fuse = ctx().opt.fuse_swap_trsm;
        // (a 65..128-row triangle needs 200 KB of shared memory per CTA: worth it only for the narrow
        //  update on the lookahead's critical path, measured)
        // wide ranges (the bulk of the trailing update, off the lookahead's critical chain): interchanges as one bandwidth
        // pass, then the whole k-row triangle in one register-tiled launch (trsm_strip.cu); 2 = also for k <= 64
        const int64_t strip = ctx().opt.trsm_strip;
        if (strip != 0 && c1 - c0 > 512 && k <= 256 && (k > 64 || strip == 2)) {
            LAIR_CHECK(swap_cols(c0, c1, k0, k0 + k, st));
            return trsm_strip_dev<T>(k, c1 - c0, at(k0, lc0), lda, at(k0, c0), lda, st);
        }
        if ((fuse == 1 && (k <= 64 || c1 - c0 <= 512)) || (fuse == 2 && c1 - c0 <= 512)) {
            const int rc = laswp_trsm_dev<T>(c1 - c0, A + c0, lda, k0, k, ipiv, at(k0, lc0), lda, st);
            if (rc != LAIR_B200_ERR_UNSUPPORTED) return rc;
        }
        if (fuse == 1 && k > 64 && k <= 256) {
            // a wide step as a chain of 64-row fused launches: each applies its own interchanges, takes
            // the contribution of the rows solved before it (prefix) and solves its own triangle --
            // 2..4 launches instead of laswp + the launch-per-block trsm recursion
            for (int64_t off = 0; off < k; off += 64) {
                const int64_t kk = (k - off) < 64 ? (k - off) : 64;
                LAIR_CHECK(laswp_trsm_dev<T>(c1 - c0, A + c0, lda, k0 + off, kk, ipiv, at(k0 + off, lc0 + off), lda, st, off));
            }
            return LAIR_B200_OK;
        }
        LAIR_CHECK(swap_cols(c0, c1, k0, k0 + k, st));                                            // laswp  (getrf.rs:270-277)
        if constexpr (sizeof(T) == 8) {
            // a tall triangle (a late column chunk catching up): one persistent dataflow solve instead
            // of the launch-per-block recursion
            // (main stream only: the solver's workspace is not shared with the lookahead stream)
            if (k > 256 && st == s && ctx().opt.trsm_dataflow >= 1) return dtrsm_ll_dev(false, k, c1 - c0, at(k0, lc0), lda, at(k0, c0), lda, st);
        }
        return trsm_lower_unit_dev<T>(k, c1 - c0, at(k0, lc0), lda, at(k0, c0), lda, st);         // trsm   (:278-283)
    }

    // Factor the w columns stored at (local) columns [c0, c0+w), whose diagonal starts at row r0:
    // rows r0..m participate, pivots land in ipiv[r0..r0+w).  On one GPU c0 == r0; under the
    // block-cyclic column distribution (mg.cu) c0 is the block's local column offset.
    int rec(int64_t r0, int64_t c0, int64_t w, cudaStream_t st) const {
        const int64_t rows = m - r0;
        const int wp = panel_max_width<T>(rows);
        if (wp <= 0) {
            set_error("getrf: %lld rows exceed the on-chip panel capacity", (long long)rows);
            return LAIR_B200_ERR_UNSUPPORTED;
        }
        if (w <= wp) return panel_dev<T>(rows, w, at(r0, c0), lda, ipiv + r0, (int32_t)r0, info, (int32_t)r0, st);
        int64_t w1 = (w / 2 + wp - 1) / wp * wp;  // left half, a multiple of the panel width
        if (w1 >= w) w1 = w - wp > 0 ? (w - 1) / wp * wp : wp;
        LAIR_CHECK(rec(r0, c0, w1, st));
        const int64_t r1 = r0 + w1, cr = c0 + w1, w2 = w - w1;
        LAIR_CHECK(swap_solve(r0, w1, c0, cr, cr + w2, st));                                      // laswp + trsm
        if (m > r1)
            LAIR_CHECK(gemm_minus_dev<T>(m - r1, w2, w1, at(r1, c0), lda, at(r0, cr), lda, at(r1, cr), lda, st));  // gemm (:289-296)
        LAIR_CHECK(rec(r1, cr, w2, st));                                                          // recurse (:297)
        return swap_cols(c0, cr, r1, r0 + w, st);                                                 // laswp left (:308-315)
    }
    int rec(int64_t j0, int64_t w, cudaStream_t st) const { return rec(j0, j0, w, st); }

    // trailing update of columns [c0, c1) with the factored block [j0, j0+jb)
    int update(int64_t j0, int64_t jb, int64_t c0, int64_t c1, cudaStream_t st) const {
        if (c1 <= c0) return LAIR_B200_OK;
        const int64_t r1 = j0 + jb;
        LAIR_CHECK(swap_solve(j0, jb, j0, c0, c1, st));
        if (r1 < m) LAIR_CHECK(gemm_minus_dev<T>(m - r1, c1 - c0, jb, at(r1, j0), lda, at(j0, c0), lda, at(r1, c0), lda, st));
        return LAIR_B200_OK;
    }

    // Paired update (run(): two block steps share one K = jp + jb trailing GEMM).  Block P = [pj0, pj0 + jp) is the
    // predecessor of block B = [j0, j0 + jb); columns [ca, cb) have P's U rows but not yet P's contribution to the rows
    // below P, and column block P already follows B's interchanges.  Brings [ca, cb) up to date with both blocks:
    // B's interchanges, the rows of B take  -= L(P) U(P)  and are solved against L11(B), the rows below take one GEMM with
    // K = jp + jb over the adjacent column blocks P and B.  Every element sees the same FMAs in the same order as two
    // separate steps (P's jp eliminations, then B's), so the result is bit-identical.
    int pair_update(int64_t pj0, int64_t jp, int64_t j0, int64_t jb, int64_t ca, int64_t cb, cudaStream_t st) const {
        if (cb <= ca) return LAIR_B200_OK;
        const int64_t r1 = j0 + jb;
        if (cb - ca <= 512) {
            // narrow (the lookahead's next block): the fused 64-row launches take P's rows as their prefix
            for (int64_t off = 0; off < jb; off += 64) {
                const int64_t kk = (jb - off) < 64 ? (jb - off) : 64;
                LAIR_CHECK(laswp_trsm_dev<T>(cb - ca, A + ca, lda, j0 + off, kk, ipiv, at(j0 + off, j0 + off), lda, st, jp + off));
            }
        } else {
            LAIR_CHECK(swap_cols(ca, cb, j0, r1, st));
            LAIR_CHECK(gemm_minus_dev<T>(jb, cb - ca, jp, at(j0, pj0), lda, at(pj0, ca), lda, at(j0, ca), lda, st));
            LAIR_CHECK(trsm_strip_dev<T>(jb, cb - ca, at(j0, j0), lda, at(j0, ca), lda, st));
        }
        if (r1 < m) LAIR_CHECK(gemm_minus_dev<T>(m - r1, cb - ca, jp + jb, at(r1, pj0), lda, at(pj0, ca), lda, at(r1, ca), lda, st));
        return LAIR_B200_OK;
    }

    // Right-looking sweep with one block of lookahead: the panel path of block k+1 (stream P,
    // high priority) runs under the bulk of block k's trailing update (stream M).
    //   M: wait panel(k) | left laswp | update(next block) -> EN | update(rest)
    //   P: wait EN | rec(next block) -> EP
    // The two streams only ever touch disjoint column ranges (P: the next block; M: the rest and
    // the already-factored left part) and disjoint ipiv ranges.
    int run() const {
        const int64_t kmin = m < n ? m : n;
        set_i32_kernel<<<1, 1, 0, s>>>(info, -1);
        LAIR_LAUNCH_CHECK();
        // outer block width: with lookahead the factorization is bound by the panel path up to
        // n ~ 8192 (narrow blocks keep it short) and by the trailing GEMM beyond (wide blocks feed
        // the DMMA kernel with a deeper K); measured on B200, profiles/r1_bench_history.md
        // The width follows the REMAINING size: while the trailing matrix is large the sweep is bound by
        // the GEMM on stream M (wide blocks = deeper K), once it is small by the panel chain on P.
        // (round 1 kept f32 at 64-wide blocks up to 8192 remaining columns; with the round-2 panel kernel both types switch at 6144)
        const int64_t fixed_nb = ctx().opt.nb;
        // (f32 switches to 256-wide blocks later: its tensor-core update is cheap, its panel chain is not; profiles/r2z_probe_tune2.jsonl)
        const int64_t t2 = (sizeof(T) == 4 && ctx().opt.nb_t2 == 10240) ? 12288 : ctx().opt.nb_t2;
        const int64_t t1 = ctx().opt.nb_t1 > 0 ? ctx().opt.nb_t1 : 6144;  // (re-measured with the fourth-generation panel: profiles/r2t_probe_tune.jsonl)
        auto pick = [&](int64_t j) {
            const int64_t rem = kmin - j;
            const int64_t v = fixed_nb > 0 ? fixed_nb : (rem > t2 ? 256 : (rem > t1 ? 128 : 64));
            return v < rem ? v : rem;
        };
        const bool look = ctx().opt.lookahead != 0 && kmin > pick(0);
        cudaStream_t M = s, P = look ? ctx().aux_stream : s;
        cudaEvent_t EP = ctx().ev[0], EN = ctx().ev[1];
        // Columns still arriving from the host (ColumnFeed): the sweep only touches columns < navail.
        // Chunk c joins -- M waits for its copy, then one update with everything factored so far --
        // early enough that the next two blocks always lie inside the joined range, and otherwise
        // at a quarter of its own offset, so the copies stay ahead of the sweep without the
        // sweep waiting for the whole matrix.
        const bool fed = feed && feed->nchunks > 1 && feed->chunk >= 512;  // first chunk holds the first two blocks
        int joined = fed ? 1 : 0;
        int64_t navail = fed ? (feed->chunk < n ? feed->chunk : n) : n;
        if (feed)  // the first chunk (or, when not sweeping chunk by chunk, every chunk) has landed
            for (int c = 0; c < (fed ? 1 : feed->nchunks); ++c) LAIR_CUDA_CHECK(cudaStreamWaitEvent(M, feed->ready[c], 0));
        auto join_due = [&](int64_t c0, int64_t nb2) -> int {
            while (fed && joined < feed->nchunks) {
                const int64_t cs = (int64_t)joined * feed->chunk;
                if (!(c0 >= kmin || cs < c0 + nb2 + 512 || c0 * ctx().opt.stream_join_div >= cs)) break;
                const int64_t ce = (cs + feed->chunk) < n ? (cs + feed->chunk) : n;
                LAIR_CUDA_CHECK(cudaStreamWaitEvent(M, feed->ready[joined], 0));
                LAIR_CHECK(update(0, c0, cs, ce, M));  // bring the late columns up to date with blocks [0, c0)
                navail = ce;
                ++joined;
            }
            return LAIR_B200_OK;
        };
        const bool chain_on_p = ctx().opt.chain_on_p != 0;
        cudaEvent_t EM = ctx().ev[2], EL = ctx().ev[3];
        // two block steps share one trailing GEMM while more than `pair_min` columns remain (0 = never)
        const int64_t pair_min = (fixed_nb == 0 && sizeof(T) == 8) ? ctx().opt.pair_k512 : 0;
        // the same for the 128- and 64-wide blocks of the later sweep (K = 256 / 128 instead of 128 / 64), while more than `pair_small` columns remain
        const int64_t pair_small = (fixed_nb == 0) ? (sizeof(T) == 8 ? ctx().opt.pair_small : ctx().opt.pair_small_f32) : 0;
        bool pend = false;
        int64_t pj0 = 0, pjb = 0;
        if (look) {
            LAIR_CUDA_CHECK(cudaEventRecord(EN, M));  // P starts after everything already queued on the caller's stream
            LAIR_CUDA_CHECK(cudaStreamWaitEvent(P, EN, 0));
            if (chain_on_p) LAIR_CUDA_CHECK(cudaEventRecord(EM, M));
        }
        int64_t jb = pick(0);
        LAIR_CHECK(rec(0, jb, P));
        for (int64_t j0 = 0; j0 < kmin; j0 += jb, jb = pick(j0)) {
            const int64_t c0 = j0 + jb;                      // first column right of the block
            const int64_t nb2 = (c0 < kmin) ? pick(c0) : 0;  // width of the next block to factor
            if (look) {
                LAIR_CUDA_CHECK(cudaEventRecord(EP, P));
                LAIR_CUDA_CHECK(cudaStreamWaitEvent(M, EP, 0));
            }
            // Rows [0, j0) are final here on M's timeline: the blocks above got their left interchanges at the end of their own
            // step, their U rows from updates M queued or P finished before EP, and later steps only touch rows >= j0.  (While
            // column chunks are still arriving their rows are not complete yet: the drain starts when the last one has joined.)
            if (drain && navail == n && drain->used < drain->nev && j0 - drain->drained >= drain->min_rows) {
                cudaEvent_t& ev = drain->ev[drain->used];
                if (!ev) LAIR_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                LAIR_CUDA_CHECK(cudaEventRecord(ev, M));
                LAIR_CUDA_CHECK(cudaStreamWaitEvent(drain->stream, ev, 0));
                LAIR_CUDA_CHECK(cudaMemcpy2DAsync((T*)drain->host + drain->drained * drain->host_rs, (size_t)drain->host_rs * sizeof(T),
                                                  at(drain->drained, 0), (size_t)lda * sizeof(T), (size_t)n * sizeof(T),
                                                  (size_t)(j0 - drain->drained), cudaMemcpyDeviceToHost, drain->stream));
                drain->drained = j0;
                ++drain->used;
            }
            const bool on_p = look && chain_on_p && kmin - j0 > ctx().opt.chain_on_p;
            if (pend) {
                // block (j0, jb) is the successor of the deferred block: columns >= c0 take both blocks at once, K = pjb + jb
                cudaStream_t X = (nb2 > 0 && on_p) ? P : M;
                if (X == P) LAIR_CUDA_CHECK(cudaStreamWaitEvent(P, EM, 0));
                LAIR_CHECK(swap_cols(pj0, j0, j0, j0 + jb, X));  // L of the deferred block follows this block's interchanges first
                if (X == P) {
                    LAIR_CUDA_CHECK(cudaEventRecord(EL, P));
                    LAIR_CUDA_CHECK(cudaStreamWaitEvent(M, EL, 0));
                }
                if (nb2 > 0) {
                    LAIR_CHECK(pair_update(pj0, pjb, j0, jb, c0, c0 + nb2, X));
                    if (X == M && look) {
                        LAIR_CUDA_CHECK(cudaEventRecord(EN, M));
                        LAIR_CUDA_CHECK(cudaStreamWaitEvent(P, EN, 0));
                    }
                    LAIR_CHECK(rec(c0, nb2, P));
                }
                LAIR_CHECK(pair_update(pj0, pjb, j0, jb, c0 + nb2, navail, M));
                LAIR_CHECK(swap_cols(0, pj0, j0, j0 + jb, M));   // the interchanges left of the deferred block
                pend = false;
                LAIR_CHECK(join_due(c0, nb2));
                if (look && chain_on_p) LAIR_CUDA_CHECK(cudaEventRecord(EM, M));
                continue;
            }
            // deferral: the next block still gets this block's full update (its panel needs it); the rest only the U rows,
            // the GEMM waits for the next block so that both share one K = 512 launch (DMMA GEMM in place at n = 65 536:
            // 32.2 TFLOP/s at K = 256, 33.6 at K = 512; profiles/r2x_probe_nb512.jsonl)
            const bool pair_here = (jb == 256 && pair_min > 0 && kmin - j0 > pair_min) ||
                                   (jb < 256 && jb >= 64 && pair_small > 0 && kmin - j0 > pair_small);
            const bool defer = pair_here && nb2 == jb && navail == n && c0 + nb2 < n && (!fed || joined == feed->nchunks);
            if (nb2 > 0) {
                if (on_p) {
                    // the whole dependent chain  panel(k) -> update(next block) -> panel(k+1)  stays on P:
                    // no cross-stream hand-over on the critical path.  P only waits for M's previous
                    // rest-update (which touched the next block's columns), long finished when the
                    // panel chain is the bottleneck.
                    LAIR_CUDA_CHECK(cudaStreamWaitEvent(P, EM, 0));
                    LAIR_CHECK(update(j0, jb, c0, c0 + nb2, P));
                } else {
                    LAIR_CHECK(update(j0, jb, c0, c0 + nb2, M));  // next block first ...
                    if (look) {
                        LAIR_CUDA_CHECK(cudaEventRecord(EN, M));
                        LAIR_CUDA_CHECK(cudaStreamWaitEvent(P, EN, 0));
                    }
                }
                LAIR_CHECK(rec(c0, nb2, P));                  // ... so its panel path can start under the rest
            }
            if (defer) {
                LAIR_CHECK(swap_solve(j0, jb, j0, c0 + nb2, navail, M));  // this block's U rows of the rest; rows below wait
                pend = true;
                pj0 = j0;
                pjb = jb;
            } else {
                LAIR_CHECK(update(j0, jb, c0 + nb2, navail, M));
            }
            LAIR_CHECK(swap_cols(0, j0, j0, j0 + jb, M));     // interchanges reach back into L (off the critical path)
            LAIR_CHECK(join_due(c0, nb2));                    // late column chunks catch up with blocks [0, c0): after
                                                              // the left interchanges, so L's rows match the pivots
            if (look && chain_on_p) LAIR_CUDA_CHECK(cudaEventRecord(EM, M));
        }
        if (look) {  // the caller's stream sees the last panel too (already implied, kept explicit)
            LAIR_CUDA_CHECK(cudaEventRecord(EP, P));
            LAIR_CUDA_CHECK(cudaStreamWaitEvent(M, EP, 0));
        }
        return LAIR_B200_OK;
    }
};

}  // namespace

template <class T>
int getrf_blocked_dev(int64_t m, int64_t n, T* d_a, int64_t lda, int32_t* d_ipiv, int32_t* d_info, cudaStream_t s,
                      const ColumnFeed* feed, RowDrain* drain) {
    LAIR_REQUIRE(m >= 0 && n >= 0 && lda >= n, "getrf: bad shape m=%lld n=%lld lda=%lld", (long long)m, (long long)n,
                 (long long)lda);
    LAIR_REQUIRE(m < (1ll << 31) && n < (1ll << 31), "getrf: dimension too large");
    if (m == 0 || n == 0) return LAIR_B200_OK;
    Factor<T> f{d_a, lda, m, n, d_ipiv, d_info, s, feed, drain};
    return f.run();
}

// Factor one block column of a (possibly column-distributed) matrix: the w columns stored at local
// columns [c0, c0+w) of d_a (m rows, leading dimension lda), diagonal starting at row r0.
template <class T>
int getrf_block_dev(int64_t m, T* d_a, int64_t lda, int64_t r0, int64_t c0, int64_t w, int32_t* d_ipiv, int32_t* d_info,
                    cudaStream_t s) {
    Factor<T> f{d_a, lda, m, c0 + w, d_ipiv, d_info, s};
    return f.rec(r0, c0, w, s);
}
template int getrf_block_dev<float>(int64_t, float*, int64_t, int64_t, int64_t, int64_t, int32_t*, int32_t*, cudaStream_t);
template int getrf_block_dev<double>(int64_t, double*, int64_t, int64_t, int64_t, int64_t, int32_t*, int32_t*, cudaStream_t);

// X = U^-1 L^-1 P B, in place in d_b (n x nrhs row-major): getrs.rs:22-36 for every column.
template <class T>
int getrs_blocked_dev(int64_t n, int64_t nrhs, const T* d_lu, int64_t lda, const int32_t* d_ipiv, T* d_b, int64_t ldb,
                      cudaStream_t s) {
    LAIR_REQUIRE(n >= 0 && nrhs >= 0 && lda >= n && ldb >= nrhs, "getrs: bad shape");
    if (n == 0 || nrhs == 0) return LAIR_B200_OK;
    // b <- P b (getrs.rs:22): all n interchanges on a tall narrow matrix collapse into one permutation
    if (ctx().opt.laswp_perm != 0 && n >= 512 && (double)n * (double)nrhs * sizeof(T) <= 256.0 * 1024 * 1024)
        LAIR_CHECK(laswp_perm_dev<T>(n, nrhs, d_b, ldb, 0, n, d_ipiv, s));
    else
        LAIR_CHECK(laswp_dev<T>(nrhs, d_b, ldb, 0, n, d_ipiv, s));
    if constexpr (sizeof(T) == 8) {
        if (ctx().opt.trsm_dataflow >= 1) {
            // flag-in-data exchange + pre-inverted diagonal blocks (trsm_ll.cu)
            LAIR_CHECK(dtrsm_ll_dev(false, n, nrhs, d_lu, lda, d_b, ldb, s));
            return dtrsm_ll_dev(true, n, nrhs, d_lu, lda, d_b, ldb, s);
        }
    }
    LAIR_CHECK(trsm_lower_unit_dev<T>(n, nrhs, d_lu, lda, d_b, ldb, s));
    return trsm_upper_dev<T>(n, nrhs, d_lu, lda, d_b, ldb, s);
}

#define INST(T)                                                                                          \
    template int getrf_blocked_dev<T>(int64_t, int64_t, T*, int64_t, int32_t*, int32_t*, cudaStream_t, const ColumnFeed*, RowDrain*); \
    template int getrs_blocked_dev<T>(int64_t, int64_t, const T*, int64_t, const int32_t*, T*, int64_t, cudaStream_t);
INST(float)
INST(double)
#undef INST

}  // namespace lair
