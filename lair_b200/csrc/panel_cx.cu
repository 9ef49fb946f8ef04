// Leaf panel of the blocked complex factorization (blocked_cx.cu): up to 8 columns, all remaining rows, factored by
// ONE thread-block cluster that keeps the panel in the shared memory of its CTAs (row slabs), exchanging per column
// one pivot candidate per CTA and the two rows of the interchange through distributed shared memory.
//
// Arithmetic = the reference's row-major loop operation for operation (src/lapack/getrf.rs:46-120 on Complex:
// iamax on |re| + |im| with first-maximum ties, src/blas/iamax.rs:6-21; reciprocal, scaled column and rank-1 update with
// separately rounded complex products, as num-complex 0.4 writes them), so a leaf returns bit for bit what the
// single-CTA kernel of small_lu.cu returns on the same panel -- the cluster only spreads the rows over up to 16 SMs
// (the single CTA is bound by one SM's path to L2: ~260 us per 8192 x 8 leaf; profiles/r1c_complex_blocked.md).
//
// Per column: local arg-max -> candidate pushed into every peer's slot array -> cluster barrier -> everyone picks the
// winner, fetches the pivot row (and the owners of rows j / p the row they receive) from the owner's shared memory ->
// cluster barrier -> owners write the interchanged rows -> block barrier -> scale + update of the CTA's own rows.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace lair {
namespace {

constexpr int PW = 8;        // panel width (storage)
constexpr int PLD = PW + 1;  // padded row: a column walk by consecutive threads is bank-conflict free (16- and 8-byte elements)
constexpr int PTHREADS = 256;
constexpr int PMAXC = 16;

template <class R>
struct Cand {
    R val;
    int idx;
    int pad;
};

template <class R>
__device__ __forceinline__ void better(R& v, int& i, R ov, int oi) {
    if (ov > v || (ov == v && oi < i)) {  // first maximum: larger value, or the same value at a lower row
        v = ov;
        i = oi;
    }
}

template <class T>
__global__ void __launch_bounds__(PTHREADS, 1)
panel_cx_kernel(T* __restrict__ A, long long lda, int M, int w, int rpc, int32_t* __restrict__ ipiv, int row_base,
                int32_t* __restrict__ info, int std_layout) {
    using O = Ops<T>;
    using R = typename O::Real;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ Cand<R> cands[PMAXC];         // slot r: the candidate of CTA r (pushed by CTA r)
    __shared__ T urow[PW], jrow[PW];
    __shared__ R red_val[PTHREADS / 32];
    __shared__ int red_idx[PTHREADS / 32];

    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    T* a = reinterpret_cast<T*>(smem_raw);   // [rpc][PLD]
    const int g0 = rank * rpc;               // first panel row of this CTA
    const int nloc = max(0, min(rpc, M - g0));

    for (int idx = tid; idx < nloc * w; idx += PTHREADS) {
        const int r = idx / w, c = idx - r * w;
        a[r * PLD + c] = A[(long long)(g0 + r) * lda + c];
    }
    cluster.sync();  // every CTA is resident before the first remote access; the slab is visible block-wide

    const int kmin = M < w ? M : w;
    int sing = -1;
    for (int j = 0; j < kmin; ++j) {
        // ---- (a) arg-max of column j over this CTA's rows >= j ----
        R bv = R(0);
        int bi = j;
        for (int r = tid; r < nloc; r += PTHREADS) {
            const int g = g0 + r;
            if (g >= j) {
                const R v = O::abs1(a[r * PLD + j]);
                if (v > bv) {  // strict: NaN never wins; rows ascend within a thread
                    bv = v;
                    bi = g;
                }
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const R ov = __shfl_xor_sync(0xffffffffu, bv, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            better(bv, bi, ov, oi);
        }
        if (lane == 0) {
            red_val[warp] = bv;
            red_idx[warp] = bi;
        }
        __syncthreads();
        if (warp == 0) {
            R v = lane < PTHREADS / 32 ? red_val[lane] : R(0);
            int i = lane < PTHREADS / 32 ? red_idx[lane] : j;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const R ov = __shfl_xor_sync(0xffffffffu, v, off);
                const int oi = __shfl_xor_sync(0xffffffffu, i, off);
                better(v, i, ov, oi);
            }
            if (lane < C) {  // push to peer `lane`
                Cand<R>* dst = cluster.map_shared_rank(&cands[rank], lane);
                dst->val = v;
                dst->idx = i;
            }
        }
        cluster.sync();
        // ---- (b) the winner; fetch the rows of the interchange from their owners ----
        R maxv = cands[0].val;
        int p = cands[0].idx;
        for (int c = 1; c < C; ++c) better(maxv, p, cands[c].val, cands[c].idx);
        if (!(maxv > R(0))) p = j;  // all-zero / all-NaN column: index 0 of the sub-column (iamax.rs:10-11)
        const int rj = j / rpc, rp = p / rpc;
        if (tid < w) {
            const T* src = cluster.map_shared_rank(a, rp);
            urow[tid] = src[(p - rp * rpc) * PLD + tid];
        } else if (tid >= 32 && tid < 32 + w && rank == rp && p != j) {
            const T* src = cluster.map_shared_rank(a, rj);
            jrow[tid - 32] = src[(j - rj * rpc) * PLD + (tid - 32)];
        }
        cluster.sync();
        // ---- (c) interchange (getrf.rs:62-71), limited to the panel's columns; the caller swaps the rest ----
        if (p != j) {
            if (rank == rj && tid < w) a[(j - g0) * PLD + tid] = urow[tid];
            if (rank == rp && tid >= 32 && tid < 32 + w) a[(p - g0) * PLD + (tid - 32)] = jrow[tid - 32];
        }
        if (rank == 0 && tid == 0) ipiv[j] = row_base + p;
        __syncthreads();
        const T pivot = urow[j];
        const bool singular = std_layout ? (maxv == R(0)) : O::is_zero(pivot);
        if (singular) {  // cluster-uniform
            sing = j;
            continue;
        }
        // ---- (d) scale and rank-1 update of this CTA's rows below j (getrf.rs:76-88) ----
        const T recip = O::recip(pivot);
        for (int r = tid; r < nloc; r += PTHREADS) {
            if (g0 + r > j) {
                T* row = a + r * PLD;
                const T l = O::mul(row[j], recip);
                row[j] = l;
                for (int k = j + 1; k < w; ++k) row[k] = O::sub(row[k], O::mul(l, urow[k]));
            }
        }
        // the next column's arg-max reads only rows this thread just updated; urow / jrow / cands are next written
        // after the next cluster barrier, which every thread reaches only after finishing (d)
    }
    __syncthreads();
    for (int idx = tid; idx < nloc * w; idx += PTHREADS) {
        const int r = idx / w, c = idx - r * w;
        A[(long long)(g0 + r) * lda + c] = a[r * PLD + c];
    }
    if (rank == 0 && tid == 0 && sing >= 0) *info = row_base + sing;  // the LAST zero-pivot step so far (getrf.rs:72-73)
    cluster.sync();  // no CTA leaves while a peer could still address its shared memory
}

}  // namespace

// Returns LAIR_B200_ERR_UNSUPPORTED (without setting an error) when the panel does not fit one cluster.
template <class T>
int panel_cx_dev(int64_t rows, int64_t w, T* d_a, int64_t lda, int32_t* d_ipiv, int32_t row_base, int32_t* d_info, bool std_layout,
                 cudaStream_t s) {
    if (w > PW || rows < 1 || w < 1) return LAIR_B200_ERR_UNSUPPORTED;
    auto kern = panel_cx_kernel<T>;
    const size_t limit = ctx().smem_optin > 8192 ? ctx().smem_optin - 8192 : 0;
    const int64_t cap = (int64_t)(limit / (PLD * sizeof(T)));  // rows per CTA
    static int max_cluster = -1;
    static uint64_t seen_epoch = 0;
    if (stale_for_context(seen_epoch)) max_cluster = -1;
    if (max_cluster < 0) {
        LAIR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit));
        max_cluster = 8;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(16);
            cfg.blockDim = dim3(PTHREADS);
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
    // enough CTAs that a thread owns at most ~2 rows, and that the slab fits
    int csize = 1;
    while (csize < max_cluster && ((rows + csize - 1) / csize > 2 * PTHREADS)) csize *= 2;
    while (csize <= max_cluster && (rows + csize - 1) / csize > cap) csize *= 2;
    if (csize > max_cluster) return LAIR_B200_ERR_UNSUPPORTED;
    const int64_t rpc = (rows + csize - 1) / csize;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(csize);
    cfg.blockDim = dim3(PTHREADS);
    cfg.dynamicSmemBytes = (size_t)rpc * PLD * sizeof(T);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csize;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    ProfScope prof(kProfPanel, s, 2.0 * (double)rows * (double)w * sizeof(T));
    LAIR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, d_a, (long long)lda, (int)rows, (int)w, (int)rpc, d_ipiv, (int)row_base, d_info,
                                       std_layout ? 1 : 0));
    LAIR_LAUNCH_CHECK();
    return LAIR_B200_OK;
}
template int panel_cx_dev<cxf>(int64_t, int64_t, cxf*, int64_t, int32_t*, int32_t, int32_t*, bool, cudaStream_t);
template int panel_cx_dev<cxd>(int64_t, int64_t, cxd*, int64_t, int32_t*, int32_t, int32_t*, bool, cudaStream_t);

}  // namespace lair
