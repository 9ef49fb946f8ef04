// Multi-RHS triangular solves for getrs, second generation (src/lapack/getrs.rs:24-36 for every
// column of B): the persistent dataflow kernel of trsm_dataflow.cu with its two per-step
// latencies removed.  Measured there (profiles/r1_bench_history.md): 1.8 ms per triangle at
// n = 8192 for ANY number of right-hand sides and either block height -- 256 sequential steps
// of ~7 us, each paying  __threadfence + flag store | flag poll | tile load | 32-step
// substitution.  Here
//   * X_i travels in a flag-in-data buffer (the LL protocol of NCCL): every double is published
//     as one 16-byte unit {lo32, epoch, hi32, epoch}; each 8-byte half is self-validating, so
//     the producer needs no fence and no flag, and the consumer's one load returns data and
//     readiness together;
//   * the 32 x 32 diagonal blocks are inverted up front by a separate, fully parallel kernel
//     (one warp per block, one column of the inverse per lane, exact substitution order), so
//     the diagonal step on the critical path is one more DMMA product instead of a 32-step
//     dependent substitution.  (The reference divides by the diagonal, getrs.rs:35; multiplying
//     by the inverse of a 32 x 32 block differs in rounding only -- the parity tests bound the
//     solution against the oracle and the residual against the oracle's.)
// Off-diagonal work is unchanged: acc -= L_ik * X_k on DMMA m8n8k4, tiles double-buffered by
// cp.async, CTA c owns row blocks c, c+G, ... and the grid fits the GPU (all CTAs become resident).
#include "common.cuh"

namespace lair {
namespace {

constexpr int RB = 32;           // rows per block
constexpr int NT = 64;           // right-hand sides per CTA
constexpr int LL_THREADS = 128;  // 4 warps: 2 (rows) x 2 (cols), warp tile 16 x 32
constexpr int LDA_S = RB + 4;    // 36 doubles: A-fragment bank = 8g + 2t
constexpr int LDX_S = NT + 4;    // 68 doubles: B-fragment bank = 8t + 2g
constexpr int UNITS = RB * NT;   // LL units (16 bytes each) per X tile

struct LLSmem {
    double a[2][RB * LDA_S];   // L / U tiles
    uint4 xl[2][UNITS];        // raw LL tiles
    double x[RB * LDX_S];      // validated, compact X tile (DMMA B operand)
    double d[RB * LDA_S];      // inverse of the diagonal block (DMMA A operand)
};

__device__ __forceinline__ void cpa16(void* smem_dst, const void* gmem_src, int bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cpa8(void* smem_dst, const void* gmem_src, int bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cpa_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void st_ll(uint4* p, double v, unsigned epoch) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};\n" ::"l"(p), "r"((unsigned)b), "r"(epoch),
                 "r"((unsigned)(b >> 32)), "r"(epoch)
                 : "memory");
}

__device__ __forceinline__ uint4 ld_ll(const uint4* p) {
    uint4 q;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(p) : "memory");
    return q;
}

// RB x RB tile of LU (rows r0.., columns k0..) into shared; rows >= rv / cols >= cv zero-filled
__device__ __forceinline__ void load_a_tile(double* __restrict__ dst, const double* __restrict__ src, long long ld_g, int rv, int cv,
                                            bool al16, int tid) {
    constexpr int CPR = RB / 2;
#pragma unroll
    for (int c = tid; c < RB * CPR; c += LL_THREADS) {
        const int r = c / CPR, col = (c % CPR) * 2;
        int valid = (r < rv) ? (cv - col) : 0;
        valid = valid < 0 ? 0 : (valid > 2 ? 2 : valid);
        const double* s = src + (long long)(r < rv ? r : 0) * ld_g + (valid > 0 ? col : 0);
        double* d = dst + r * LDA_S + col;
        if (al16) {
            cpa16(d, s, valid * 8);
        } else {
            cpa8(d, s, valid > 0 ? 8 : 0);
            cpa8(d + 1, valid > 1 ? s + 1 : s, valid > 1 ? 8 : 0);
        }
    }
}

// ---- inverse of every RB x RB diagonal block (unit-lower part, or upper part) ----------------
// One warp per block, lane j computes column j of the inverse by the substitution the reference
// would run on the unit vector e_j (same order of operations as getrs.rs:24-36).
template <bool UPPER>
__global__ void __launch_bounds__(128)
tri_inv_blocks_kernel(const double* __restrict__ LU, long long lda, int n, double* __restrict__ inv) {
    __shared__ double sT[4][RB * (RB + 1)];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int blk = blockIdx.x * 4 + warp;
    const int nblk = (n + RB - 1) / RB;
    if (blk >= nblk) return;
    const int r0 = blk * RB;
    const int rv = (n - r0) < RB ? (n - r0) : RB;
    double* T = sT[warp];
    for (int r = 0; r < RB; ++r) {  // lane = column: coalesced rows
        double v = (r == lane) ? 1.0 : 0.0;
        if (r < rv && lane < rv) {
            const bool keep = UPPER ? (lane >= r) : (lane < r);
            if (keep) v = LU[(long long)(r0 + r) * lda + r0 + lane];
        }
        T[r * (RB + 1) + lane] = v;
    }
    __syncwarp();
    double x[RB];
    if (!UPPER) {
#pragma unroll
        for (int i = 0; i < RB; ++i) {
            double s = (i == lane) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < i; ++k) s -= T[i * (RB + 1) + k] * x[k];
            x[i] = s;
        }
    } else {
#pragma unroll
        for (int i = RB - 1; i >= 0; --i) {
            double s = (i == lane) ? 1.0 : 0.0;
#pragma unroll
            for (int k = RB - 1; k > i; --k) s -= T[i * (RB + 1) + k] * x[k];
            x[i] = s / T[i * (RB + 1) + i];
        }
    }
    double* out = inv + (size_t)blk * RB * RB;
#pragma unroll
    for (int i = 0; i < RB; ++i) out[i * RB + lane] = x[i];  // element (i, lane): coalesced rows
}

template <bool UPPER>
__global__ void __launch_bounds__(LL_THREADS)
dtrsm_ll_kernel(const double* __restrict__ LU, long long lda, int n, double* __restrict__ B, long long ldb, int nrhs,
                const double* __restrict__ inv, uint4* __restrict__ ll, unsigned epoch, int* __restrict__ err) {
    extern __shared__ __align__(16) unsigned char ll_smem_raw[];
    LLSmem& sm = *reinterpret_cast<LLSmem*>(ll_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 1, wn = warp & 1;
    const int nblk = (n + RB - 1) / RB;
    const int ct = blockIdx.y;  // column tile
    const int c0 = ct * NT;
    const int cv = (nrhs - c0) < NT ? (nrhs - c0) : NT;
    uint4* llt = ll + (size_t)ct * nblk * UNITS;
    const bool al_lu = ((lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(LU) & 15) == 0);

    // CTA x owns steps x, x+G, ...: while the chain of dependent diagonal steps is elsewhere, a CTA
    // works ahead on its next block.  The grid never exceeds what an otherwise idle GPU keeps resident
    // (host side), so every CTA eventually runs: kernels of the lookahead stream that share the GPU
    // finish on their own and only delay residency.
    for (int s = blockIdx.x; s < nblk; s += gridDim.x) {
        const int blk = UPPER ? (nblk - 1 - s) : s;
        const int r0 = blk * RB;
        const int rv = (n - r0) < RB ? (n - r0) : RB;
        auto dep_block = [&](int d) { return UPPER ? (nblk - 1 - d) : d; };
        auto issue = [&](int d, int buf, bool with_a, bool with_x) {
            const int kb = dep_block(d);
            if (with_a) {
                const int k0 = kb * RB;
                const int kv = (n - k0) < RB ? (n - k0) : RB;
                load_a_tile(sm.a[buf], LU + (long long)r0 * lda + k0, lda, rv, kv, al_lu, tid);
            }
            if (with_x) {
                const uint4* src = llt + (size_t)kb * UNITS;
#pragma unroll
                for (int u = tid; u < UNITS; u += LL_THREADS) cpa16(&sm.xl[buf][u], src + u, 16);
            }
        };

        // accumulators = this block's tile of B; (row g, cols 2t, 2t+1) of every 8x8 tile
        double acc[2][4][2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = wm * 16 + i * 8 + g;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = wn * 32 + j * 8 + 2 * t;
                const double* p = B + (long long)(r0 + r) * ldb + c0 + c;
                acc[i][j][0] = (r < rv && c < cv) ? p[0] : 0.0;
                acc[i][j][1] = (r < rv && c + 1 < cv) ? p[1] : 0.0;
            }
        }
        // inverse of the diagonal block (needed last; in flight during the updates)
        {
            const double* src = inv + (size_t)blk * RB * RB;
#pragma unroll
            for (int c = tid; c < RB * (RB / 2); c += LL_THREADS) {
                const int r = c / (RB / 2), col = (c % (RB / 2)) * 2;
                cpa16(&sm.d[r * LDA_S + col], src + r * RB + col, 16);
            }
        }
        if (s > 0) issue(0, 0, true, s > 1);

        for (int d = 0; d < s; ++d) {
            const int buf = d & 1;
            // ---- tile d: validate the flag-in-data units and compact them into sm.x ----
            // Early dependencies were published long ago: their units were prefetched by cp.async and
            // are checked in shared memory.  The LAST dependency is the chain predecessor, still being
            // produced: its units are polled straight from L2, all of a thread's loads in flight at once.
            cpa_wait_all();
            __syncthreads();  // raw tile visible; everyone is done with sm.x and buffer buf^1 (tile d-1)
            bool ok = false;
            if (d + 1 < s) {
                ok = true;
#pragma unroll 4
                for (int u = tid; u < UNITS; u += LL_THREADS) {
                    const uint4 q = sm.xl[buf][u];
                    ok = ok && (q.y == epoch) && (q.w == epoch);
                    const int k = u / NT, c = u % NT;
                    sm.x[k * LDX_S + c] = __longlong_as_double((long long)(((unsigned long long)q.z << 32) | q.x));
                }
                ok = __syncthreads_and(ok);
            }
            if (!ok) {
                constexpr int PER = UNITS / LL_THREADS;
                const uint4* src = llt + (size_t)dep_block(d) * UNITS;
                uint4 q[PER];
                int spins = 0;
                for (;;) {
#pragma unroll
                    for (int i = 0; i < PER; ++i) q[i] = ld_ll(src + tid + i * LL_THREADS);
                    bool all = true;
#pragma unroll
                    for (int i = 0; i < PER; ++i) all = all && (q[i].y == epoch) && (q[i].w == epoch);
                    if (all) break;
                    if (++spins > (1 << 22)) {
                        atomicExch(err, 2);
                        break;
                    }
                }
#pragma unroll
                for (int i = 0; i < PER; ++i) {
                    const int u = tid + i * LL_THREADS;
                    sm.x[(u / NT) * LDX_S + (u % NT)] = __longlong_as_double((long long)(((unsigned long long)q[i].z << 32) | q[i].x));
                }
                __syncthreads();
            }
            if (d + 1 < s) issue(d + 1, buf ^ 1, true, d + 2 < s);  // X units only for dependencies checked in shared memory
            const double* a_s = sm.a[buf];
#pragma unroll
            for (int kk = 0; kk < RB / 4; ++kk) {
                double af[2], bf[4];
#pragma unroll
                for (int i = 0; i < 2; ++i) af[i] = -a_s[(wm * 16 + i * 8 + g) * LDA_S + kk * 4 + t];
#pragma unroll
                for (int j = 0; j < 4; ++j) bf[j] = sm.x[(kk * 4 + t) * LDX_S + wn * 32 + j * 8 + g];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
            }
        }
        cpa_wait_all();
        __syncthreads();  // all warps done with sm.x; the inverse block has landed

        // ---- diagonal step: X_i = inv(T_ii) * acc, one more DMMA product ----
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = wm * 16 + i * 8 + g;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = wn * 32 + j * 8 + 2 * t;
                sm.x[r * LDX_S + c] = acc[i][j][0];
                sm.x[r * LDX_S + c + 1] = acc[i][j][1];
            }
        }
        __syncthreads();
        double out[2][4][2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) out[i][j][0] = out[i][j][1] = 0.0;
#pragma unroll
        for (int kk = 0; kk < RB / 4; ++kk) {
            double af[2], bf[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) af[i] = sm.d[(wm * 16 + i * 8 + g) * LDA_S + kk * 4 + t];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = sm.x[(kk * 4 + t) * LDX_S + wn * 32 + j * 8 + g];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(out[i][j][0], out[i][j][1], af[i], bf[j]);
        }
        // ---- publish X_i: flag-in-data units for the consumers, plain values into B ----
        uint4* dst = llt + (size_t)blk * UNITS;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int r = wm * 16 + i * 8 + g;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = wn * 32 + j * 8 + 2 * t;
                st_ll(dst + r * NT + c, out[i][j][0], epoch);
                st_ll(dst + r * NT + c + 1, out[i][j][1], epoch);
                double* p = B + (long long)(r0 + r) * ldb + c0 + c;
                if (r < rv && c < cv) p[0] = out[i][j][0];
                if (r < rv && c + 1 < cv) p[1] = out[i][j][1];
            }
        }
        __syncthreads();  // sm.x / sm.d are reused by this CTA's next block
    }
}

struct LLState {
    void* buf = nullptr;  // [inverse blocks | LL units | err]
    size_t inv_doubles = 0, units = 0;
    unsigned epoch = 0;
    int grid_cap = -1;
};
LLState g_ll;
void reset_ll_state() {
    if (g_ll.buf) cudaFree(g_ll.buf);
    g_ll = LLState();
}
ResetHook g_ll_hook(reset_ll_state);

}  // namespace

// X = T^-1 B in place for T = unit-lower (upper = false) or upper (upper = true) n x n in d_lu.
int dtrsm_ll_dev(bool upper, int64_t n, int64_t nrhs, const double* d_lu, int64_t lda, double* d_b, int64_t ldb, cudaStream_t s) {
    if (n == 0 || nrhs == 0) return LAIR_B200_OK;
    LAIR_REQUIRE(n < (1ll << 30) && nrhs < (1ll << 30), "trsm: dimension too large");
    const int nblk = (int)((n + RB - 1) / RB);
    const int ntile = (int)((nrhs + NT - 1) / NT);
    const size_t need_inv = (size_t)nblk * RB * RB;
    const size_t need_units = (size_t)nblk * ntile * UNITS;
    LLState& st = g_ll;
    if (st.inv_doubles < need_inv || st.units < need_units) {
        if (st.buf) {
            LAIR_CUDA_CHECK(cudaDeviceSynchronize());
            LAIR_CUDA_CHECK(cudaFree(st.buf));
            st.buf = nullptr;
        }
        const size_t inv_d = need_inv > st.inv_doubles ? need_inv : st.inv_doubles;
        const size_t un = need_units > st.units ? need_units : st.units;
        const size_t bytes = inv_d * sizeof(double) + un * sizeof(uint4) + 64;
        LAIR_CUDA_CHECK(cudaMalloc(&st.buf, bytes));
        LAIR_CUDA_CHECK(cudaMemset(st.buf, 0, bytes));  // epoch 0 is never used: every unit starts invalid
        st.inv_doubles = inv_d;
        st.units = un;
        st.epoch = 0;
    }
    double* d_inv = reinterpret_cast<double*>(st.buf);
    uint4* d_units = reinterpret_cast<uint4*>(d_inv + st.inv_doubles);
    int* d_err = ctx().d_fault;  // a wait that times out raises the context's fault word (check_fault)
    constexpr size_t kSmem = sizeof(LLSmem);
    if (st.grid_cap < 0) {
        LAIR_CUDA_CHECK(cudaFuncSetAttribute(dtrsm_ll_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
        LAIR_CUDA_CHECK(cudaFuncSetAttribute(dtrsm_ll_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
        int per_sm = 0, per_sm_u = 0;
        LAIR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dtrsm_ll_kernel<false>, LL_THREADS, kSmem));
        LAIR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_u, dtrsm_ll_kernel<true>, LL_THREADS, kSmem));
        if (per_sm_u < per_sm) per_sm = per_sm_u;
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 2) per_sm = 2;
        st.grid_cap = per_sm * ctx().sm_count;
    }
    if (++st.epoch == 0) {  // wrapped: stale units could look valid
        LAIR_CUDA_CHECK(cudaDeviceSynchronize());
        LAIR_CUDA_CHECK(cudaMemset(d_units, 0, st.units * sizeof(uint4)));
        st.epoch = 1;
    }
    // all CTAs must be able to be resident together (they wait on each other): grid.x * grid.y <= capacity
    int gx = st.grid_cap / ntile;
    if (gx < 1) {
        set_error("trsm: %d right-hand-side tiles exceed the co-resident CTA capacity", ntile);
        return LAIR_B200_ERR_UNSUPPORTED;
    }
    if (gx > nblk) gx = nblk;
    ProfScope prof(kProfTrsm, s, (double)n * (double)n * (double)nrhs);
    const unsigned inv_grid = (unsigned)((nblk + 3) / 4);
    if (upper)
        tri_inv_blocks_kernel<true><<<inv_grid, 128, 0, s>>>(d_lu, (long long)lda, (int)n, d_inv);
    else
        tri_inv_blocks_kernel<false><<<inv_grid, 128, 0, s>>>(d_lu, (long long)lda, (int)n, d_inv);
    LAIR_LAUNCH_CHECK();
    dim3 grid((unsigned)gx, (unsigned)ntile);
    if (upper)
        dtrsm_ll_kernel<true><<<grid, LL_THREADS, kSmem, s>>>(d_lu, (long long)lda, (int)n, d_b, (long long)ldb, (int)nrhs, d_inv,
                                                              d_units, st.epoch, d_err);
    else
        dtrsm_ll_kernel<false><<<grid, LL_THREADS, kSmem, s>>>(d_lu, (long long)lda, (int)n, d_b, (long long)ldb, (int)nrhs, d_inv,
                                                               d_units, st.epoch, d_err);
    LAIR_LAUNCH_CHECK();
    return LAIR_B200_OK;
}

}  // namespace lair
