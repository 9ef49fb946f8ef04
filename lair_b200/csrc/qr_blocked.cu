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
//   3. trailing update C := (I - V T^T V^T) C as three products on the DMMA / FFMA GEMM kernels (gemm_minus_dev: C -= A B):
//      W = V^T C  (W = 0; W -= NVt C),  W := T^T W (32-row triangle, one small kernel),  C -= Vp W.
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
        NVt[(long long)k * ldt + r] = -v;
    }
}

// W (QB x ncols, ld ldw) := T^T W, T upper triangular QB x QB (row-major ld QB)
template <class R>
__global__ void __launch_bounds__(128)
qr_trmm_tt_kernel(const R* __restrict__ T, R* __restrict__ W, long long ldw, int ncols) {
    __shared__ R Ts[QB][QB + 1];
    for (int idx = threadIdx.x; idx < QB * QB; idx += 128) Ts[idx / QB][idx % QB] = T[idx];
    __syncthreads();
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c >= ncols) return;
    R wv[QB];
#pragma unroll
    for (int k = 0; k < QB; ++k) wv[k] = W[(long long)k * ldw + c];
#pragma unroll
    for (int i = QB - 1; i >= 0; --i) {
        R s = R(0);
#pragma unroll
        for (int k = 0; k <= i; ++k) s += Ts[k][i] * wv[k];
        W[(long long)i * ldw + c] = s;
    }
}

struct QrWork {
    void* p = nullptr;
    size_t bytes = 0;
};
int qr_workspace(size_t bytes, void** out, cudaStream_t s) {
    static QrWork wk;
    if (wk.bytes < bytes) {
        if (wk.p) {
            LAIR_CUDA_CHECK(cudaStreamSynchronize(s));
            LAIR_CUDA_CHECK(cudaFree(wk.p));
            wk = QrWork{};
        }
        LAIR_CUDA_CHECK(cudaMalloc(&wk.p, bytes + bytes / 4 + 256));
        wk.bytes = bytes + bytes / 4 + 256;
    }
    *out = wk.p;
    return LAIR_B200_OK;
}

template <class R>
int qr_panel_dev(int64_t rows, int64_t w, R* d_a, int64_t lda, R* d_tau, R* d_T, cudaStream_t s) {
    auto kern = qr_panel_kernel<R>;
    const size_t limit = ctx().smem_optin > 20480 ? ctx().smem_optin - 20480 : 0;
    const int64_t cap = (int64_t)(limit / (QLD * sizeof(R)));
    static int max_cluster = -1;
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

// Returns LAIR_B200_ERR_UNSUPPORTED (nothing done, no error set) when the first panel does not fit one cluster.
template <class R>
int geqrf_blocked_dev(int64_t m, int64_t n, R* d_a, int64_t lda, R* d_tau, cudaStream_t s) {
    const int64_t k = m < n ? m : n;
    if (k == 0) return LAIR_B200_OK;
    // workspace: T (QB x QB) | Vp (m x QB) | NVt (QB x ldt) | W (QB x n)
    const int64_t ldt = (m + 3) / 4 * 4, ldw = (n + 3) / 4 * 4;
    const size_t off_T = 0, off_V = 4096 * sizeof(R), off_N = off_V + (size_t)m * QB * sizeof(R), off_W = off_N + (size_t)QB * ldt * sizeof(R);
    const size_t total = off_W + (size_t)QB * ldw * sizeof(R);
    {   // capacity check before anything is modified
        const size_t limit = ctx().smem_optin > 20480 ? ctx().smem_optin - 20480 : 0;
        const int64_t cap = (int64_t)(limit / (QLD * sizeof(R)));
        if ((m + QMAXC - 1) / QMAXC > cap) return LAIR_B200_ERR_UNSUPPORTED;
    }
    void* ws = nullptr;
    LAIR_CHECK(qr_workspace(total, &ws, s));
    R* dT = reinterpret_cast<R*>((char*)ws + off_T);
    R* dV = reinterpret_cast<R*>((char*)ws + off_V);
    R* dN = reinterpret_cast<R*>((char*)ws + off_N);
    R* dW = reinterpret_cast<R*>((char*)ws + off_W);
    for (int64_t j0 = 0; j0 < k; j0 += QB) {
        const int64_t jb = (k - j0) < QB ? (k - j0) : QB;
        const int64_t rows = m - j0;
        R* ajj = d_a + j0 * lda + j0;
        // the panel may be wider than jb when the matrix is wide and this is the last block: only jb reflectors exist
        const int rc = qr_panel_dev<R>(rows, jb, ajj, lda, d_tau + j0, dT, s);
        if (rc != LAIR_B200_OK) {
            if (rc == LAIR_B200_ERR_UNSUPPORTED) set_error("geqrf: panel of %lld rows does not fit one cluster", (long long)rows);
            return rc == LAIR_B200_ERR_UNSUPPORTED ? LAIR_B200_ERR_CUDA : rc;
        }
        const int64_t nc = n - j0 - jb;
        if (nc <= 0) continue;
        const unsigned pb = (unsigned)std::min<int64_t>((rows * QB + 255) / 256, (int64_t)ctx().sm_count * 8);
        qr_pack_kernel<R><<<pb, 256, 0, s>>>(ajj, (long long)lda, (int)rows, (int)jb, dV, dN, (long long)ldt);
        LAIR_LAUNCH_CHECK();
        LAIR_CUDA_CHECK(cudaMemsetAsync(dW, 0, (size_t)QB * ldw * sizeof(R), s));
        R* c = ajj + jb;
        LAIR_CHECK(gemm_minus_dev<R>(QB, nc, rows, dN, ldt, c, lda, dW, ldw, s));       // W = V^T C
        qr_trmm_tt_kernel<R><<<(unsigned)((nc + 127) / 128), 128, 0, s>>>(dT, dW, (long long)ldw, (int)nc);  // W := T^T W
        LAIR_LAUNCH_CHECK();
        LAIR_CHECK(gemm_minus_dev<R>(rows, nc, QB, dV, QB, dW, ldw, c, lda, s));        // C -= V W
    }
    return LAIR_B200_OK;
}
template int geqrf_blocked_dev<float>(int64_t, int64_t, float*, int64_t, float*, cudaStream_t);
template int geqrf_blocked_dev<double>(int64_t, int64_t, double*, int64_t, double*, cudaStream_t);

}  // namespace lair
