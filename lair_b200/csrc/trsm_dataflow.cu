// Multi-RHS triangular solves for getrs (src/lapack/getrs.rs:24-36 for every column of B) as
// ONE persistent dataflow kernel per triangle instead of a launch-per-block recursion.
//
// B (n x nrhs, row-major) is cut into row blocks of RB = 64 (or 32) rows.  CTA c owns blocks c, c+G, ...
// (all CTAs co-resident).  For its block i it keeps the RB x 64 tile of B in DMMA
// accumulators, and for every earlier block k (later block, for the upper solve) it
//   waits for X_k to be published (a flag word in global memory, acquire load),
//   streams the RB x RB tile of L (or U) and the RB x 64 tile X_k into shared memory
//   (cp.async, double buffered) and accumulates  acc -= L_ik * X_k  with DMMA m8n8k4,
// then solves its RB x RB diagonal block by substitution, 32 rows at a time (one thread per
// right-hand side, the 32 unknowns in registers, true divide by the diagonal for U as the
// reference does; between the two halves of a 64-row block a rank-32 update by all threads),
// writes X_i, fences, and publishes flag[i].  The critical path is one flag round trip plus
// one small update and one diagonal solve per block instead of ~4 kernel launches.
// Roofline: latency-bound at nrhs = 64 (2 n^2 nrhs flops); DMMA does all off-diagonal flops.
#include "common.cuh"

namespace lair {
namespace {

constexpr int HB = 32;          // rows of one substitution unit (the diagonal block is solved HB rows at a time)
constexpr int NT = 64;          // right-hand sides per CTA
constexpr int DF_THREADS = 128;  // 4 warps: 2 (rows) x 2 (cols), warp tile RB/2 x 32
constexpr int LDX_S = NT + 4;   // 68 doubles: B-fragment bank = 8t + 2g
// RB (template parameter) = rows per block: 32, or 64 (half as many sequential steps on the
// critical path; the 64 x 64 diagonal block is solved as 32 | rank-32 update | 32).
template <int RB> struct DfCfg {
    static constexpr int LDA_S = RB + 4;  // A-fragment bank = 8g + 2t (36 or 68 doubles)
    static constexpr int LDD = RB + 1;
    static constexpr int MI = RB / 16;    // 8-row DMMA tiles per warp
    static constexpr size_t smem = (size_t)(2 * RB * (LDA_S + LDX_S) + RB * LDD) * sizeof(double);
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src, int bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

// ROWS x COLS tile (doubles) from global (row pitch ld_g) into shared (row pitch ld_s);
// rows >= rv / cols >= cv are zero-filled.  `al16`: 16-byte chunks are aligned.
template <int ROWS, int COLS>
__device__ __forceinline__ void load_tile(double* __restrict__ dst, int ld_s, const double* __restrict__ src, long long ld_g, int rv,
                                          int cv, bool al16, int tid) {
    constexpr int CPR = COLS / 2;
#pragma unroll
    for (int c = tid; c < ROWS * CPR; c += DF_THREADS) {
        const int r = c / CPR, col = (c % CPR) * 2;
        int valid = (r < rv) ? (cv - col) : 0;
        valid = valid < 0 ? 0 : (valid > 2 ? 2 : valid);
        const double* s = src + (long long)(r < rv ? r : 0) * ld_g + (valid > 0 ? col : 0);
        double* d = dst + r * ld_s + col;
        if (al16) {
            cp_async16(d, s, valid * 8);
        } else {
            cp_async8(d, s, valid > 0 ? 8 : 0);
            cp_async8(d + 1, valid > 1 ? s + 1 : s, valid > 1 ? 8 : 0);
        }
    }
}

template <bool UPPER, int RB>
__global__ void __launch_bounds__(DF_THREADS)
dtrsm_dataflow_kernel(const double* __restrict__ LU, long long lda, int n, double* __restrict__ B, long long ldb, int nrhs,
                      unsigned* __restrict__ flags, unsigned epoch, int* __restrict__ err) {
    constexpr int LDA_S = DfCfg<RB>::LDA_S, LDD = DfCfg<RB>::LDD, MI = DfCfg<RB>::MI;
    extern __shared__ __align__(16) unsigned char df_smem[];
    double(*sA)[RB * LDA_S] = reinterpret_cast<double(*)[RB * LDA_S]>(df_smem);                              // [2]
    double(*sX)[RB * LDX_S] = reinterpret_cast<double(*)[RB * LDX_S]>(df_smem + 2 * RB * LDA_S * sizeof(double));  // [2]
    double* sD = reinterpret_cast<double*>(df_smem + 2 * RB * (LDA_S + LDX_S) * sizeof(double));             // diagonal block
    __shared__ int s_known_slot;
    int* s_known = &s_known_slot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    const int nblk = (n + RB - 1) / RB;
    const int ct = blockIdx.y;  // column tile
    const int c0 = ct * NT;
    const int cv = (nrhs - c0) < NT ? (nrhs - c0) : NT;
    unsigned* fl = flags + (size_t)ct * nblk;
    const bool al_lu = ((lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(LU) & 15) == 0);
    const bool al_b = ((ldb & 1) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && ((c0 & 1) == 0);

    for (int s = blockIdx.x; s < nblk; s += gridDim.x) {
        const int blk = UPPER ? (nblk - 1 - s) : s;
        const int r0 = blk * RB;
        const int rv = (n - r0) < RB ? (n - r0) : RB;

        // accumulators = this block's tile of B; (row g, cols 2t, 2t+1) of every 8x8 tile
        double acc[MI][4][2];
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            const int r = wm * (RB / 2) + i * 8 + g;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = wn * 32 + j * 8 + 2 * t;
                const double* p = B + (long long)(r0 + r) * ldb + c0 + c;
                acc[i][j][0] = (r < rv && c < cv) ? p[0] : 0.0;
                acc[i][j][1] = (r < rv && c + 1 < cv) ? p[1] : 0.0;
            }
        }
        // diagonal block (needed last; loaded first so it is in flight during the updates)
        for (int idx = tid; idx < RB * RB; idx += DF_THREADS) {
            const int r = idx / RB, c = idx % RB;
            double v = (r == c) ? 1.0 : 0.0;
            if (r < rv && c < rv) {
                const bool keep = UPPER ? (c >= r) : (c < r);
                if (keep) v = LU[(long long)(r0 + r) * lda + r0 + c];
            }
            sD[r * LDD + c] = v;
        }

        // ---- off-diagonal updates, dependency d = 0 .. s-1, software pipelined by one ----
        // `known` = number of leading dependencies known to be published.  It is refreshed by ONE
        // warp reading 32 flags at a time, so a late block (all of whose early dependencies were
        // published long ago) streams its tiles without a flag round trip per tile; only the
        // dependencies still being produced are polled.
        auto dep_block = [&](int d) { return UPPER ? (nblk - 1 - d) : d; };
        int known = 0;
        auto poll = [&](int d) {  // all threads; one L2 round trip
            if (warp == 0) {
                const int idx = d + lane;
                const bool ok = (idx < s) && (ld_acquire(fl + dep_block(idx)) == epoch);
                const unsigned m = __ballot_sync(0xffffffffu, ok);
                if (lane == 0) *s_known = d + (m == 0xffffffffu ? 32 : (__ffs(~m) - 1));
            }
            __syncthreads();
            known = *s_known;
            __syncthreads();
        };
        auto wait_ready = [&](int d) {
            int spins = 0;
            while (known <= d) {
                poll(d);
                if (++spins > (1 << 22)) {
                    if (tid == 0) atomicExch(err, 2);
                    break;
                }
            }
        };
        // the L (or U) tile does not depend on the flag: it can be in flight while the flag is awaited
        auto issue_a = [&](int d, int buf) {
            const int k0 = dep_block(d) * RB;
            const int kv = (n - k0) < RB ? (n - k0) : RB;
            load_tile<RB, RB>(sA[buf], LDA_S, LU + (long long)r0 * lda + k0, lda, rv, kv, al_lu, tid);
        };
        auto issue_x = [&](int d, int buf) {
            const int k0 = dep_block(d) * RB;
            const int kv = (n - k0) < RB ? (n - k0) : RB;
            load_tile<RB, NT>(sX[buf], LDX_S, B + (long long)k0 * ldb + c0, ldb, kv, cv, al_b, tid);
            cp_async_commit();
        };
        if (s > 0) {
            issue_a(0, 0);
            wait_ready(0);
            issue_x(0, 0);
        }
        for (int d = 0; d < s; ++d) {
            const int buf = d & 1;
            cp_async_wait_all();
            __syncthreads();  // tile d visible to all; everyone is done with buffer buf^1 (tile d-1)
            // prefetch tile d+1 now if its producer has already published; otherwise start its L tile,
            // compute first and block afterwards (keeps this tile's flops off the critical path)
            bool issued = false;
            if (d + 1 < s) {
                issue_a(d + 1, buf ^ 1);
                if (known <= d + 1) poll(d + 1);
                if (known > d + 1) {
                    issue_x(d + 1, buf ^ 1);
                    issued = true;
                }
            }
            const double* a_s = sA[buf];
            const double* x_s = sX[buf];
#pragma unroll
            for (int kk = 0; kk < RB / 4; ++kk) {
                double af[MI], bf[4];
#pragma unroll
                for (int i = 0; i < MI; ++i) af[i] = -a_s[(wm * (RB / 2) + i * 8 + g) * LDA_S + kk * 4 + t];
#pragma unroll
                for (int j = 0; j < 4; ++j) bf[j] = x_s[(kk * 4 + t) * LDX_S + wn * 32 + j * 8 + g];
#pragma unroll
                for (int i = 0; i < MI; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
            if (d + 1 < s && !issued) {
                wait_ready(d + 1);
                issue_x(d + 1, buf ^ 1);
            }
        }
        __syncthreads();  // all warps done with the tile buffers; sD complete

        // ---- diagonal solve: accumulators -> shared; HB rows at a time, one thread per right-hand side,
        //      then a rank-HB update of the block's remaining rows by all threads ----
        double* xs = sX[0];
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            const int r = wm * (RB / 2) + i * 8 + g;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = wn * 32 + j * 8 + 2 * t;
                xs[r * LDX_S + c] = acc[i][j][0];
                xs[r * LDX_S + c + 1] = acc[i][j][1];
            }
        }
        __syncthreads();
        constexpr int NS = RB / HB;
#pragma unroll
        for (int q = 0; q < NS; ++q) {
            const int sb = UPPER ? (NS - 1 - q) : q;  // sub-block solved in this round
            const int o = sb * HB;                    // its first row / column inside the block
            if (tid < NT) {
                double b[HB];
#pragma unroll
                for (int i = 0; i < HB; ++i) b[i] = xs[(o + i) * LDX_S + tid];
                if (!UPPER) {
#pragma unroll
                    for (int kk = 0; kk < HB - 1; ++kk) {
                        const double bk = b[kk];
#pragma unroll
                        for (int i = kk + 1; i < HB; ++i) b[i] -= sD[(o + i) * LDD + o + kk] * bk;
                    }
                } else {
#pragma unroll
                    for (int i = HB - 1; i >= 0; --i) {
                        b[i] = b[i] / sD[(o + i) * LDD + o + i];  // true divide (getrs.rs:35)
                        const double bi = b[i];
#pragma unroll
                        for (int r = 0; r < i; ++r) b[r] -= sD[(o + r) * LDD + o + i] * bi;
                    }
                }
#pragma unroll
                for (int i = 0; i < HB; ++i) xs[(o + i) * LDX_S + tid] = b[i];
            }
            if (q + 1 < NS) {
                __syncthreads();
                // rows of the sub-blocks still to solve -= D[rows, o..o+HB) * X[o..o+HB)
                const int rem_rows = (NS - 1 - q) * HB;
                const int row_lo = UPPER ? 0 : o + HB;
                const int c = tid % NT, rpart = tid / NT;  // DF_THREADS / NT row groups
                constexpr int RG = DF_THREADS / NT;
                double xk[HB];
#pragma unroll
                for (int k = 0; k < HB; ++k) xk[k] = xs[(o + k) * LDX_S + c];
                for (int r = rpart; r < rem_rows; r += RG) {
                    const int row = row_lo + r;
                    double v = xs[row * LDX_S + c];
#pragma unroll
                    for (int k = 0; k < HB; ++k) v -= sD[row * LDD + o + k] * xk[k];
                    xs[row * LDX_S + c] = v;
                }
                __syncthreads();
            }
        }
        __syncthreads();
        // ---- publish X_i ----
        for (int idx = tid; idx < RB * NT; idx += DF_THREADS) {
            const int r = idx / NT, c = idx % NT;
            if (r < rv && c < cv) B[(long long)(r0 + r) * ldb + c0 + c] = xs[r * LDX_S + c];
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release(fl + blk, epoch);
    }
}

struct DataflowState {
    unsigned* flags = nullptr;
    size_t cap = 0;      // flag words
    unsigned epoch = 0;
};
DataflowState g_df;

template <int RB>
int launch_dataflow(bool upper, int64_t n, int64_t nrhs, const double* d_lu, int64_t lda, double* d_b, int64_t ldb, cudaStream_t s) {
    const int nblk = (int)((n + RB - 1) / RB);
    const int ntile = (int)((nrhs + NT - 1) / NT);
    const size_t need = (size_t)nblk * ntile;
    DataflowState& st = g_df;
    if (st.cap < need + 1) {
        if (st.flags) {
            LAIR_CUDA_CHECK(cudaDeviceSynchronize());
            LAIR_CUDA_CHECK(cudaFree(st.flags));
        }
        size_t cap = need * 2 + 1024;
        LAIR_CUDA_CHECK(cudaMalloc(&st.flags, (cap + 1) * sizeof(unsigned)));
        LAIR_CUDA_CHECK(cudaMemset(st.flags, 0, (cap + 1) * sizeof(unsigned)));
        st.cap = cap;
        st.epoch = 0;
    }
    constexpr size_t kSmem = DfCfg<RB>::smem;
    static int grid_cap = -1;
    if (grid_cap < 0) {
        LAIR_CUDA_CHECK(cudaFuncSetAttribute(dtrsm_dataflow_kernel<false, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
        LAIR_CUDA_CHECK(cudaFuncSetAttribute(dtrsm_dataflow_kernel<true, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
        int per_sm = 0;
        LAIR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dtrsm_dataflow_kernel<false, RB>, DF_THREADS, kSmem));
        int per_sm_u = 0;
        LAIR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_u, dtrsm_dataflow_kernel<true, RB>, DF_THREADS, kSmem));
        if (per_sm_u < per_sm) per_sm = per_sm_u;
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 2) per_sm = 2;  // leave room: the kernel is latency-, not occupancy-bound
        grid_cap = per_sm * ctx().sm_count;
    }
    if (++st.epoch == 0) {  // wrapped: flags may hold stale equal values
        LAIR_CUDA_CHECK(cudaDeviceSynchronize());
        LAIR_CUDA_CHECK(cudaMemset(st.flags, 0, st.cap * sizeof(unsigned)));
        st.epoch = 1;
    }
    // all CTAs must be co-resident (they wait on each other): grid.x * grid.y <= capacity
    int gx = grid_cap / ntile;
    if (gx < 1) {
        set_error("trsm: %d right-hand-side tiles exceed the co-resident CTA capacity", ntile);
        return LAIR_B200_ERR_UNSUPPORTED;
    }
    if (gx > nblk) gx = nblk;
    dim3 grid((unsigned)gx, (unsigned)ntile);
    ProfScope prof(kProfTrsm, s, (double)n * (double)n * (double)nrhs);
    if (upper)
        dtrsm_dataflow_kernel<true, RB><<<grid, DF_THREADS, kSmem, s>>>(d_lu, (long long)lda, (int)n, d_b, (long long)ldb, (int)nrhs,
                                                                        st.flags, st.epoch, ctx().d_fault);
    else
        dtrsm_dataflow_kernel<false, RB><<<grid, DF_THREADS, kSmem, s>>>(d_lu, (long long)lda, (int)n, d_b, (long long)ldb, (int)nrhs,
                                                                         st.flags, st.epoch, ctx().d_fault);
    LAIR_LAUNCH_CHECK();
    return LAIR_B200_OK;
}

}  // namespace

// X = T^-1 B in place for T = unit-lower (UPPER=false) or upper (UPPER=true) n x n in d_lu.
// Option trsm_rb picks the row-block height (32 default, 64; measured equally fast: the chain of
// dependent steps costs the same either way -- trsm_ll.cu attacks the per-step latency instead).
int dtrsm_dataflow_dev(bool upper, int64_t n, int64_t nrhs, const double* d_lu, int64_t lda, double* d_b, int64_t ldb,
                       cudaStream_t s) {
    if (n == 0 || nrhs == 0) return LAIR_B200_OK;
    LAIR_REQUIRE(n < (1ll << 30) && nrhs < (1ll << 30), "trsm: dimension too large");
    if (ctx().opt.trsm_rb == 64) return launch_dataflow<64>(upper, n, nrhs, d_lu, lda, d_b, ldb, s);
    return launch_dataflow<32>(upper, n, nrhs, d_lu, lda, d_b, ldb, s);
}

}  // namespace lair
