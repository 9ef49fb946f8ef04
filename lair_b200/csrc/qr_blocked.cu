// Blocked Householder QR for f32 / f64 on the panel + trailing-GEMM machinery of the LU path.
//
// Per block of QB = 32 columns (compact WY, the shape of src/lapack/larft.rs + larfb.rs around geqrf.rs:9-30):
//   1. qr_panel_kernel: ONE thread-block cluster keeps the rows x 32 panel in the shared memory of its CTAs (row slabs) and
//      runs the reference's loop on it -- larfg (src/lapack/larfg.rs:9-42) then larf::left on the panel's remaining columns
//      (src/lapack/larf.rs:10-54) with the reference's operation structure (separately rounded products and sums) --
//      exchanging per column the partial |x|^2 and the partial dot products v^T a_k through distributed shared memory
//      (two cluster barriers per column).  The same dot products against the columns LEFT of j are V^T v_j, from which
//      the kernel builds the triangular factor T (larft's forward / columnwise recurrence, src/lapack/larft.rs) on the fly.
//   2. pack: Vp = V with its unit diagonal and zeros above (rows x 32), NVt = -V^T (32 x rows).
//   3. trailing update C := (I - V T^T V^T) C:  W = V^T C by a split-row SIMT kernel (a reduction over the rows: partial
//      products per 512-row chunk, summed in order inside the next kernel),  W := T^T W (32-row triangle),  C -= Vp W on
//      the DMMA / FFMA GEMM kernel of the LU path (gemm_minus_dev).
// tau, R and the reflectors come out in the reference's (LAPACK's) storage; results agree with the unblocked loop to
// rounding (tests/test_gpu_qr.py: blocked vs unblocked vs oracle).
#include <cooperative_groups.h>

#include <algorithm>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace lair {
namespace {

constexpr int QB = 32;          // block / panel width
constexpr int QLD = QB + 1;     // padded slab row
constexpr int QTHREADS = 256;
constexpr int QROWG = QTHREADS / QB;  // row groups of the column-parallel phases
constexpr int QMAXC = 16;

template <class R> struct Rn;
template <> struct Rn<float> {
    __device__ static float mul(float a, float b) { return __fmul_rn(a, b); }
    __device__ static float add(float a, float b) { return __fadd_rn(a, b); }
    __device__ static float sub(float a, float b) { return __fsub_rn(a, b); }
    __device__ static float div(float a, float b) { return __fdiv_rn(a, b); }
    __device__ static float eps() { return 5.9604644775390625e-08f; }
    __device__ static float sfmin() { return 1.17549435082228750797e-38f; }
};
template <> struct Rn<double> {
    __device__ static double mul(double a, double b) { return __dmul_rn(a, b); }
    __device__ static double add(double a, double b) { return __dadd_rn(a, b); }
    __device__ static double sub(double a, double b) { return __dsub_rn(a, b); }
    __device__ static double div(double a, double b) { return __ddiv_rn(a, b); }
    __device__ static double eps() { return 1.1102230246251565404e-16; }
    __device__ static double sfmin() { return 2.2250738585072013831e-308; }
};

// rows x w panel (w <= 32) at A; tau[0..w), T (w x w upper triangular, row-major ld QB; strictly lower part zero).
template <class R>
__global__ void __launch_bounds__(QTHREADS, 1)
qr_panel_kernel(R* __restrict__ A, long long lda, int M, int w, int rpc, R* __restrict__ tau_out, R* __restrict__ T_out) {
    using N = Rn<R>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ R slot_part[QMAXC];        // per-CTA partial |x|^2 (slot r pushed by CTA r)
    __shared__ R slot_alpha;              // a[j][j], pushed by the owner of row j
    __shared__ R slot_dot[QMAXC][QB];     // per-CTA partial dot products v^T a_k
    __shared__ R part[QROWG][QB];
    __shared__ R red[QTHREADS / 32];
    __shared__ R Ts[QB][QB + 1];
    __shared__ R zs[QB];

    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid % QB, ty = tid / QB;
    R* a = reinterpret_cast<R*>(smem_raw);  // [rpc][QLD]
    const int g0 = rank * rpc;
    const int nloc = max(0, min(rpc, M - g0));

    for (int idx = tid; idx < nloc * w; idx += QTHREADS) {
        const int r = idx / w, c = idx - r * w;
        a[r * QLD + c] = A[(long long)(g0 + r) * lda + c];
    }
    for (int idx = tid; idx < QB * (QB + 1); idx += QTHREADS) (&Ts[0][0])[idx] = R(0);
    cluster.sync();

    const int kmin = M < w ? M : w;
    for (int j = 0; j < kmin; ++j) {
        const int rj = j / rpc;
        // cluster-wide sum over rows > j of a[r][j]^2 (parallel order), plus alpha = a[j][j] from its owner
        auto norm_sq = [&]() -> R {
            R s = R(0);
            for (int r = tid; r < nloc; r += QTHREADS)
                if (g0 + r > j) s = N::add(s, N::mul(a[r * QLD + j], a[r * QLD + j]));
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            if (lane == 0) red[warp] = s;
            __syncthreads();
            if (warp == 0) {
                R t = R(0);
                for (int q = 0; q < QTHREADS / 32; ++q) t += red[q];
                if (lane < C) *cluster.map_shared_rank(&slot_part[rank], lane) = t;
                if (rank == rj && lane < C) *cluster.map_shared_rank(&slot_alpha, lane) = a[(j - g0) * QLD + j];
            }
            cluster.sync();
            R t = R(0);
            for (int c = 0; c < C; ++c) t += slot_part[c];
            return t;
        };
        R x_norm = sqrt(norm_sq());
        R alpha = slot_alpha;
        R tau, scale, beta;
        bool identity = false;
        if (x_norm == R(0)) {  // larfg.rs:15-17 (real scalar: im == 0): H = I
            identity = true;
            tau = R(0);
            scale = R(1);
            beta = alpha;
        } else {
            beta = -copysign(sqrt(N::add(N::add(N::mul(alpha, alpha), R(0)), N::mul(x_norm, x_norm))), alpha);  // lapy3(re, 0, |x|)
            const R safe_min = N::sfmin() / N::eps();
            int knt = 0;
            if (fabs(beta) < safe_min) {  // cluster-uniform: every thread holds the same beta
                const R rsm = R(1) / safe_min;
                for (;;) {
                    ++knt;
                    for (int r = tid; r < nloc; r += QTHREADS)
                        if (g0 + r > j) a[r * QLD + j] = N::mul(a[r * QLD + j], rsm);
                    beta *= rsm;
                    alpha = N::mul(alpha, rsm);
                    if (fabs(beta) >= safe_min || knt >= 20) break;
                }
                cluster.sync();  // everyone has read slot_part / slot_alpha of the first round
                const R keep = alpha;
                x_norm = sqrt(norm_sq());
                alpha = keep;    // slot_alpha still holds the unscaled diagonal
                beta = -copysign(N::add(N::mul(alpha, alpha), N::mul(x_norm, x_norm)), alpha);  // literal (larfg.rs:33)
            }
            tau = N::div(N::sub(beta, alpha), beta);
            scale = N::div(R(1), N::sub(alpha, beta));
            for (int k = 0; k < knt; ++k) beta *= safe_min;
        }
        // scale x, store beta / tau
        if (!identity) {
            for (int r = tid; r < nloc; r += QTHREADS) {
                const int g = g0 + r;
                if (g > j) a[r * QLD + j] = N::mul(a[r * QLD + j], scale);
                else if (g == j) a[r * QLD + j] = beta;
            }
        }
        if (tid == 0) {
            if (rank == 0) tau_out[j] = tau;
            Ts[j][j] = tau;
        }
        __syncthreads();
        if (identity) {          // tau == 0: larf::left returns at once (larf.rs:16-18); T's column j stays zero (larft)
            cluster.sync();      // peers may still be reading this round's slots: the next round's pushes must wait
            continue;
        }
        // ---- dot products of v_j = [1; x] with every other column k over rows >= j:  k > j -> w_k,  k < j -> (V^T v_j)_k ----
        {
            R acc = R(0);
            if (tx < w && tx != j) {
                for (int r = ty; r < nloc; r += QROWG) {
                    const int g = g0 + r;
                    if (g > j) acc = N::add(acc, N::mul(a[r * QLD + tx], a[r * QLD + j]));
                    else if (g == j) acc = N::add(acc, a[r * QLD + tx]);  // v[j] = 1 (and for k < j this entry is R's, not V's:
                                                                          // V[j][k] for k < j IS the stored a[j][k] -- row j > k)
                }
            }
            part[ty][tx] = acc;
            __syncthreads();
            if (ty == 0) {
                R s = part[0][tx];
#pragma unroll
                for (int g = 1; g < QROWG; ++g) s = N::add(s, part[g][tx]);
                for (int c = 0; c < C; ++c) *cluster.map_shared_rank(&slot_dot[rank][tx], c) = s;
            }
            cluster.sync();
        }
        R dotk = R(0);
        for (int c = 0; c < C; ++c) dotk = N::add(dotk, slot_dot[c][tx]);
        // ---- rank-1 update of the columns right of j: a[r][k] += (-tau v[r]) w[k]  (gerc, src/blas/gerc.rs:8-34) ----
        if (tx > j && tx < w) {
            const R nt = -tau;
            for (int r = ty; r < nloc; r += QROWG) {
                const int g = g0 + r;
                if (g >= j) {
                    const R v = g == j ? R(1) : a[r * QLD + j];
                    a[r * QLD + tx] = N::add(a[r * QLD + tx], N::mul(N::mul(nt, v), dotk));
                }
            }
        }
        // ---- T[0..j, j] = -tau T[0..j, 0..j] (V^T v_j)[0..j]  (larft, forward columnwise); every CTA keeps its own copy ----
        if (ty == 0 && tx < j) zs[tx] = dotk;
        __syncthreads();
        if (ty == 0 && tx < j) {
            R s = R(0);
            for (int k = tx; k < j; ++k) s = N::add(s, N::mul(Ts[tx][k], zs[k]));
            Ts[tx][j] = N::mul(-tau, s);
        }
        __syncthreads();
    }
    for (int idx = tid; idx < nloc * w; idx += QTHREADS) {
        const int r = idx / w, c = idx - r * w;
        A[(long long)(g0 + r) * lda + c] = a[r * QLD + c];
    }
    if (rank == 0)
        for (int idx = tid; idx < QB * QB; idx += QTHREADS) T_out[idx] = Ts[idx / QB][idx % QB];
    cluster.sync();
}

// Vp (rows x QB, ld QB): V with unit diagonal, zeros above and right of column w; NVt (QB x rows, ld ldt) = -Vp^T
template <class R>
__global__ void qr_pack_kernel(const R* __restrict__ A, long long lda, int rows, int w, R* __restrict__ Vp, R* __restrict__ NVt, long long ldt) {
    const long long total = (long long)rows * QB;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(idx / QB), k = (int)(idx - (long long)r * QB);
        R v = R(0);
        if (k < w) v = r < k ? R(0) : (r == k ? R(1) : A[(long long)r * lda + k]);
        Vp[idx] = v;
        if (NVt) NVt[(long long)k * ldt + r] = -v;
    }
}

// Partial products of W = V^T C: CTA (x, y) takes 128 columns of C and the row chunk y (VT_RC rows) and writes the
// 32 x 128 partial sum to Wp[y] (ld ldw).  A reduction over the rows is badly served by an output-tiled GEMM (M = 32,
// K = rows: 5 TFLOP/s on the DMMA kernel), so this is a SIMT kernel with 4 x 8 register tiles: per staged row 6 shared
// loads feed 32 FMAs, and the (column tile) x (row chunk) grid fills the GPU even for narrow trailing blocks.
// The partials are summed in chunk order by qr_trmm_tt_kernel: deterministic, no atomics.
constexpr int VT_BN = 128, VT_RC = 512, VT_KS = 16, VT_THREADS = 128;
// column j (0..7) of thread tj (0..15) inside the 128-column tile: 16-byte groups interleaved across the threads, so a
// quarter-warp's 16-byte shared loads fall in distinct banks (8 consecutive columns per thread would be 4-way conflicted)
template <class R>
__device__ __forceinline__ constexpr int vt_col(int tj, int j) {
    constexpr int VEC = 16 / (int)sizeof(R);
    return (j / VEC) * (16 * VEC) + tj * VEC + (j % VEC);
}
template <class R>
__global__ void __launch_bounds__(VT_THREADS)
qr_vtc_kernel(const R* __restrict__ Vp /* rows x QB */, const R* __restrict__ Cm, long long ldc, int rows, int ncols,
              R* __restrict__ Wp, long long ldw, long long part_stride, int rc = VT_RC) {
    __shared__ __align__(16) R Vs[VT_KS][QB];
    __shared__ __align__(16) R Cs[VT_KS][VT_BN];
    const int tid = threadIdx.x;
    const int ti = tid / 16, tj = tid % 16;  // V columns 4 ti .. +4, C columns 8 tj .. +8
    const int c0 = blockIdx.x * VT_BN;
    const int r0 = blockIdx.y * rc;
    const int rend = min(rows, r0 + rc);
    R acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = R(0);
    for (int rb = r0; rb < rend; rb += VT_KS) {
        __syncthreads();
#pragma unroll
        for (int q = 0; q < VT_KS * QB / VT_THREADS; ++q) {
            const int idx = tid + q * VT_THREADS;
            const int r = idx / QB, k = idx % QB;
            Vs[r][k] = (rb + r < rend) ? Vp[(long long)(rb + r) * QB + k] : R(0);
        }
#pragma unroll
        for (int q = 0; q < VT_KS * VT_BN / VT_THREADS; ++q) {
            const int idx = tid + q * VT_THREADS;
            const int r = idx / VT_BN, c = idx % VT_BN;
            Cs[r][c] = (rb + r < rend && c0 + c < ncols) ? Cm[(long long)(rb + r) * ldc + c0 + c] : R(0);
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < VT_KS; ++r) {
            R v[4], c[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = Vs[r][ti * 4 + i];
#pragma unroll
            for (int j = 0; j < 8; ++j) c[j] = Cs[r][vt_col<R>(tj, j)];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fma(v[i], c[j], acc[i][j]);
        }
    }
    R* out = Wp + (long long)blockIdx.y * part_stride;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = c0 + vt_col<R>(tj, j);
            if (c < ncols) out[(long long)(ti * 4 + i) * ldw + c] = acc[i][j];
        }
}

// W (QB x ncols, ld ldw) := T^T (sum of the nparts partial products Wp[p], ld ldp), T upper triangular QB x QB (row-major ld QB)
template <class R, bool TRANS = true>
__global__ void __launch_bounds__(128)
qr_trmm_tt_kernel(const R* __restrict__ T, const R* __restrict__ Wp, long long part_stride, long long ldp, int nparts, R* __restrict__ W, long long ldw,
                  int ncols) {
    // CTA = 32 columns x 4 groups of 8 rows: thread (c, g) sums the partial products of rows 8g .. 8g+8 of its column (in
    // chunk order), the sums meet in shared memory, then the same thread forms 8 rows of the triangular product.
    __shared__ R Ts[QB][QB + 1];
    __shared__ R ws[QB][QB + 1];
    for (int idx = threadIdx.x; idx < QB * QB; idx += 128) Ts[idx / QB][idx % QB] = T[idx];
    const int cl = threadIdx.x % QB, g = threadIdx.x / QB;
    const int c = blockIdx.x * QB + cl;
    R acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = R(0);
    if (c < ncols) {
#pragma unroll 4
        for (int p = 0; p < nparts; ++p) {
            const R* src = Wp + (long long)p * part_stride + (long long)(8 * g) * ldp + c;
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] += src[(long long)q * ldp];
        }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) ws[8 * g + q][cl] = acc[q];
    __syncthreads();
    if (c >= ncols) return;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int i = 8 * g + q;
        R s = R(0);
        if constexpr (TRANS) {  // (T^T w)[i] = sum_{k <= i} T[k][i] w[k]
            for (int k = 0; k <= i; ++k) s += Ts[k][i] * ws[k][cl];
        } else {                // (T w)[i] = sum_{k >= i} T[i][k] w[k]   (accumulating Q = H_0 H_1 ... applies T itself)
            for (int k = i; k < QB; ++k) s += Ts[i][k] * ws[k][cl];
        }
        W[(long long)i * ldw + c] = s;
    }
}

// T (QB x QB, row-major) from tau and the Gram matrix G = V^T V, given as nparts partial products (32 x 32, ld QB each):
// T[j][j] = tau_j, T[0..j, j] = -tau_j T[0..j, 0..j] G[0..j, j]  (larft, forward columnwise; src/lapack/larft.rs).
template <class R>
__global__ void __launch_bounds__(QB * QB)
qr_build_t_kernel(const R* __restrict__ Gp, long long part_stride, int nparts, const R* __restrict__ tau, int w, R* __restrict__ T) {
    __shared__ R G[QB][QB + 1];
    __shared__ R Ts[QB][QB + 1];
    const int i = threadIdx.x / QB, k = threadIdx.x % QB;
    R g = R(0);
    for (int p = 0; p < nparts; ++p) g += Gp[(long long)p * part_stride + i * QB + k];
    G[i][k] = g;
    Ts[i][k] = R(0);
    __syncthreads();
    for (int j = 0; j < w; ++j) {
        const R t = tau[j];
        if (i == 0 && k < j) {  // row k of column j
            R s = R(0);
            for (int q = k; q < j; ++q) s += Ts[k][q] * G[q][j];
            Ts[k][j] = -t * s;
        }
        if (i == 0 && k == j) Ts[j][j] = t;
        __syncthreads();
    }
    T[i * QB + k] = Ts[i][k];
}

int qr_workspace(size_t bytes, void** out, cudaStream_t s) { return ensure_work(Context::kWorkQr, bytes, out, s); }

template <class R>
int qr_panel_dev(int64_t rows, int64_t w, R* d_a, int64_t lda, R* d_tau, R* d_T, cudaStream_t s) {
    auto kern = qr_panel_kernel<R>;
    const size_t limit = ctx().smem_optin > 20480 ? ctx().smem_optin - 20480 : 0;
    const int64_t cap = (int64_t)(limit / (QLD * sizeof(R)));
    static int max_cluster = -1;
    static uint64_t seen_epoch = 0;
    if (stale_for_context(seen_epoch)) max_cluster = -1;
    if (max_cluster < 0) {
        LAIR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
        max_cluster = 8;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(16);
            cfg.blockDim = dim3(QTHREADS);
            cfg.dynamicSmemBytes = limit;
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
    int csize = 1;
    while (csize < max_cluster && ((rows + csize - 1) / csize > 2 * QTHREADS)) csize *= 2;
    while (csize <= max_cluster && (rows + csize - 1) / csize > cap) csize *= 2;
    if (csize > max_cluster) return LAIR_B200_ERR_UNSUPPORTED;
    const int64_t rpc = (rows + csize - 1) / csize;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(csize);
    cfg.blockDim = dim3(QTHREADS);
    cfg.dynamicSmemBytes = (size_t)rpc * QLD * sizeof(R);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csize;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    ProfScope prof(kProfPanel, s, 2.0 * (double)rows * (double)w * sizeof(R));
    LAIR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, d_a, (long long)lda, (int)rows, (int)w, (int)rpc, d_tau, d_T));
    LAIR_LAUNCH_CHECK();
    return LAIR_B200_OK;
}

}  // namespace

template <class R>
int geqrf_blocked_dev(int64_t m, int64_t n, R* d_a, int64_t lda, R* d_tau, cudaStream_t s) {
    const int64_t k = m < n ? m : n;
    if (k == 0) return LAIR_B200_OK;
    // workspace: T (2 x QB x QB) | Vp (m x QB) | (spare, QB x ldt) | W (QB x n) | Wp (nparts x QB x n) | V2, Gram partials (fallback panel)
    const int64_t ldt = (m + 3) / 4 * 4, ldw = (n + 3) / 4 * 4;
    const int64_t max_parts = (m + VT_RC - 1) / VT_RC;
    const size_t off_T = 0, off_V = 4096 * sizeof(R), off_N = off_V + (size_t)m * QB * sizeof(R), off_W = off_N + (size_t)QB * ldt * sizeof(R);
    const size_t off_P = off_W + (size_t)QB * ldw * sizeof(R);
    const size_t narrow_parts = (size_t)((m + 127) / 128) * QB * QB;  // the lookahead's narrow updates: 128-row chunks, 32 columns
    const size_t off_V2 = off_P + std::max((size_t)max_parts * QB * ldw, narrow_parts) * sizeof(R);  // fallback panel: its own V ...
    const size_t off_P2 = off_V2 + (size_t)m * QB * sizeof(R);                                        // ... and Gram partials
    const size_t total = off_P2 + (size_t)max_parts * QB * QB * sizeof(R);
    void* ws = nullptr;
    LAIR_CHECK(qr_workspace(total, &ws, s));
    R* dT[2] = {reinterpret_cast<R*>((char*)ws + off_T), reinterpret_cast<R*>((char*)ws + off_T) + 2048};
    R* dV = reinterpret_cast<R*>((char*)ws + off_V);
    R* dW = reinterpret_cast<R*>((char*)ws + off_W);
    R* dP = reinterpret_cast<R*>((char*)ws + off_P);
    // One block of lookahead, as in the LU sweep (blocked.cu): the panel of block b+1 (stream P, high priority, one
    // cluster = 16 SMs) runs under the bulk of block b's trailing update (stream M).
    //   M: wait panel(b) | pack | update(next block's columns) -> EN | update(rest)
    //   P: wait EN | panel(b+1) -> EP
    // P only touches the next block's columns, its tau range and the other T buffer.
    const bool look = ctx().opt.lookahead != 0 && k > QB;
    cudaStream_t M = s, P = look ? ctx().aux_stream : s;
    cudaEvent_t EP = ctx().ev[0], EN = ctx().ev[1];
    R* dV2 = reinterpret_cast<R*>((char*)ws + off_V2);
    R* dP2 = reinterpret_cast<R*>((char*)ws + off_P2);
    auto panel = [&](int64_t j0, int64_t jb, int buf, cudaStream_t st) -> int {
        const int64_t rows = m - j0;
        R* ajj = d_a + j0 * lda + j0;
        const int rc = qr_panel_dev<R>(rows, jb, ajj, lda, d_tau + j0, dT[buf], st);
        if (rc != LAIR_B200_ERR_UNSUPPORTED) return rc;
        // A panel too tall for the shared memory of one cluster: the one-reflector loop on the panel's columns (its
        // reflector application splits the rows over CTAs, qr.cu), then T from tau and V^T V as in qr_q_blocked_dev.
        LAIR_CHECK(geqrf_unblocked_dev<R>(rows, jb, ajj, lda, d_tau + j0, st));
        const unsigned pb = (unsigned)std::min<int64_t>((rows * QB + 255) / 256, (int64_t)ctx().sm_count * 8);
        qr_pack_kernel<R><<<pb, 256, 0, st>>>(ajj, (long long)lda, (int)rows, (int)jb, dV2, (R*)nullptr, 0);
        LAIR_LAUNCH_CHECK();
        const int nparts = (int)((rows + VT_RC - 1) / VT_RC);
        qr_vtc_kernel<R><<<dim3(1, (unsigned)nparts), VT_THREADS, 0, st>>>(dV2, dV2, (long long)QB, (int)rows, QB, dP2, (long long)QB, (long long)QB * QB);
        LAIR_LAUNCH_CHECK();
        qr_build_t_kernel<R><<<1, QB * QB, 0, st>>>(dP2, (long long)QB * QB, nparts, d_tau + j0, (int)jb, dT[buf]);
        LAIR_LAUNCH_CHECK();
        return LAIR_B200_OK;
    };
    // C (rows x ncols at c) := (I - V T^T V^T) C with the packed reflectors of the current block
    auto update = [&](int64_t rows, R* c, int64_t ncols, const R* T) -> int {
        if (ncols <= 0) return LAIR_B200_OK;
        // a narrow update (the lookahead's next block) sits on the panel -> update -> panel chain: smaller row chunks
        // put more CTAs on it (its partial buffer is 32 columns wide, so more chunks still fit)
        const bool narrow = ncols <= QB;
        const int rc = narrow ? 128 : VT_RC;
        const int nparts = (int)((rows + rc - 1) / rc);
        const long long ldp = narrow ? QB : ldw, pstride = (long long)QB * ldp;
        {   // W = V^T C as partial products over row chunks, summed inside the T^T kernel
            ProfScope prof(kProfTrsm, M, 2.0 * QB * (double)ncols * (double)rows);
            dim3 grid((unsigned)((ncols + VT_BN - 1) / VT_BN), (unsigned)nparts);
            qr_vtc_kernel<R><<<grid, VT_THREADS, 0, M>>>(dV, c, (long long)lda, (int)rows, (int)ncols, dP, ldp, pstride, rc);
            LAIR_LAUNCH_CHECK();
        }
        qr_trmm_tt_kernel<R><<<(unsigned)((ncols + QB - 1) / QB), 128, 0, M>>>(T, dP, pstride, ldp, nparts, dW, (long long)ldw, (int)ncols);  // W := T^T W
        LAIR_LAUNCH_CHECK();
        return gemm_minus_dev<R>(rows, ncols, QB, dV, QB, dW, ldw, c, lda, M);              // C -= V W
    };
    if (look) {
        LAIR_CUDA_CHECK(cudaEventRecord(EN, M));  // P starts after everything already queued on the caller's stream
        LAIR_CUDA_CHECK(cudaStreamWaitEvent(P, EN, 0));
    }
    LAIR_CHECK(panel(0, k < QB ? k : QB, 0, P));
    int buf = 0;
    for (int64_t j0 = 0; j0 < k; j0 += QB, buf ^= 1) {
        const int64_t jb = (k - j0) < QB ? (k - j0) : QB;
        const int64_t rows = m - j0;
        R* ajj = d_a + j0 * lda + j0;
        const int64_t c0 = j0 + jb;                                  // first column right of the block
        const int64_t nb2 = c0 < k ? ((k - c0) < QB ? (k - c0) : QB) : 0;  // width of the next block
        if (look) {
            LAIR_CUDA_CHECK(cudaEventRecord(EP, P));
            LAIR_CUDA_CHECK(cudaStreamWaitEvent(M, EP, 0));
        }
        if (c0 >= n) break;
        const unsigned pb = (unsigned)std::min<int64_t>((rows * QB + 255) / 256, (int64_t)ctx().sm_count * 8);
        qr_pack_kernel<R><<<pb, 256, 0, M>>>(ajj, (long long)lda, (int)rows, (int)jb, dV, (R*)nullptr, 0);
        LAIR_LAUNCH_CHECK();
        if (nb2 > 0) {
            LAIR_CHECK(update(rows, ajj + jb, nb2, dT[buf]));        // the next block's columns first ...
            if (look) {
                LAIR_CUDA_CHECK(cudaEventRecord(EN, M));
                LAIR_CUDA_CHECK(cudaStreamWaitEvent(P, EN, 0));
            }
            LAIR_CHECK(panel(c0, nb2, buf ^ 1, P));                   // ... so its panel can start under the rest
        }
        LAIR_CHECK(update(rows, ajj + jb + nb2, n - c0 - nb2, dT[buf]));
    }
    if (look) {  // the caller's stream sees the last panel too
        LAIR_CUDA_CHECK(cudaEventRecord(EP, P));
        LAIR_CUDA_CHECK(cudaStreamWaitEvent(M, EP, 0));
    }
    return LAIR_B200_OK;
}
// qr::Factorized::q (qr.rs:27-59) by blocks: Q = H_0 H_1 ... H_{k-1} applied to the identity from the last block
// backwards, each block as I - V T V^T on the trailing square (the standard blocked ungqr; same Q to rounding as the
// reference's reflector-at-a-time loop).  T is rebuilt per block from tau and V^T V (split-row kernel + larft recurrence).
template <class R>
__global__ void q_identity_kernel(R* __restrict__ Q, long long ldq, int m) {
    const long long total = (long long)m * m;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(idx / m), c = (int)(idx - (long long)r * m);
        Q[(long long)r * ldq + c] = r == c ? R(1) : R(0);
    }
}

template <class R>
int qr_q_blocked_dev(int64_t m, int64_t n, const R* d_qr, int64_t ldqr, const R* d_tau, R* d_q, int64_t ldq, cudaStream_t s) {
    const int64_t k = m < n ? m : n;
    const int64_t ldw = (m + 3) / 4 * 4;
    const int64_t max_parts = (m + VT_RC - 1) / VT_RC;
    // workspace: T | Vp (m x QB) | W (QB x m) | Wp (parts x QB x m)
    const size_t off_T = 0, off_V = 4096 * sizeof(R), off_W = off_V + (size_t)m * QB * sizeof(R);
    const size_t off_P = off_W + (size_t)QB * ldw * sizeof(R);
    const size_t total = off_P + (size_t)max_parts * QB * ldw * sizeof(R);
    void* ws = nullptr;
    LAIR_CHECK(qr_workspace(total, &ws, s));
    R* dT = reinterpret_cast<R*>((char*)ws + off_T);
    R* dV = reinterpret_cast<R*>((char*)ws + off_V);
    R* dW = reinterpret_cast<R*>((char*)ws + off_W);
    R* dP = reinterpret_cast<R*>((char*)ws + off_P);
    const unsigned ib = (unsigned)std::min<int64_t>((m * m + 255) / 256, (int64_t)ctx().sm_count * 8);
    q_identity_kernel<R><<<ib, 256, 0, s>>>(d_q, (long long)ldq, (int)m);
    LAIR_LAUNCH_CHECK();
    const int64_t nblocks = (k + QB - 1) / QB;
    for (int64_t b = nblocks - 1; b >= 0; --b) {
        const int64_t j0 = b * QB;
        const int64_t jb = (k - j0) < QB ? (k - j0) : QB;
        const int64_t rows = m - j0, nc = m - j0;
        const int nparts = (int)((rows + VT_RC - 1) / VT_RC);
        const unsigned pb = (unsigned)std::min<int64_t>((rows * QB + 255) / 256, (int64_t)ctx().sm_count * 8);
        qr_pack_kernel<R><<<pb, 256, 0, s>>>(d_qr + j0 * ldqr + j0, (long long)ldqr, (int)rows, (int)jb, dV, (R*)nullptr, 0);
        LAIR_LAUNCH_CHECK();
        // T from tau and V^T V
        qr_vtc_kernel<R><<<dim3(1, (unsigned)nparts), VT_THREADS, 0, s>>>(dV, dV, (long long)QB, (int)rows, QB, dP, (long long)QB, (long long)QB * QB);
        LAIR_LAUNCH_CHECK();
        qr_build_t_kernel<R><<<1, QB * QB, 0, s>>>(dP, (long long)QB * QB, nparts, d_tau + j0, (int)jb, dT);
        LAIR_LAUNCH_CHECK();
        // Q_sub := (I - V T V^T) Q_sub
        R* qs = d_q + j0 * ldq + j0;
        {
            ProfScope prof(kProfTrsm, s, 2.0 * QB * (double)nc * (double)rows);
            qr_vtc_kernel<R><<<dim3((unsigned)((nc + VT_BN - 1) / VT_BN), (unsigned)nparts), VT_THREADS, 0, s>>>(dV, qs, (long long)ldq, (int)rows, (int)nc, dP, (long long)ldw,
                                                                                                            (long long)QB * ldw);
            LAIR_LAUNCH_CHECK();
        }
        qr_trmm_tt_kernel<R, false><<<(unsigned)((nc + QB - 1) / QB), 128, 0, s>>>(dT, dP, (long long)QB * ldw, (long long)ldw, nparts, dW, (long long)ldw, (int)nc);
        LAIR_LAUNCH_CHECK();
        LAIR_CHECK(gemm_minus_dev<R>(rows, nc, QB, dV, QB, dW, ldw, qs, ldq, s));
    }
    return LAIR_B200_OK;
}
template int qr_q_blocked_dev<float>(int64_t, int64_t, const float*, int64_t, const float*, float*, int64_t, cudaStream_t);
template int qr_q_blocked_dev<double>(int64_t, int64_t, const double*, int64_t, const double*, double*, int64_t, cudaStream_t);

template int geqrf_blocked_dev<float>(int64_t, int64_t, float*, int64_t, float*, cudaStream_t);
template int geqrf_blocked_dev<double>(int64_t, int64_t, double*, int64_t, double*, cudaStream_t);

}  // namespace lair
