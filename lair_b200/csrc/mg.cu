// Multi-GPU layer: one large LU on a 1-D block-cyclic COLUMN distribution, one process per GPU.
//
// Global block column j (width nb) lives on rank j mod P as local block j / P, so the pivot
// search, the panel factorization and every row interchange are local to a rank (rows are never
// split).  The only exchange of the whole factorization is, per block column, an NCCL broadcast
// (NVLink / NVSwitch) of the factored panel (rows j*nb.. x nb, packed) and its nb pivots from the
// owner to everyone; each rank then runs laswp + trsm + DMMA gemm on its own columns.
// Lookahead: the owner of block k+1 updates that block first, factors it on the high-priority
// stream and broadcasts it while all ranks are still inside the trailing update of block k.
//
// NCCL is loaded with dlopen at mg_init time, so the single-GPU library has no NCCL dependency;
// the communicator is created from an ncclUniqueId the host layer distributes (torch.distributed
// or any other out-of-band channel): lair_b200_mg_unique_id -> lair_b200_mg_init.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <vector>

#include "common.cuh"

namespace lair {
namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, ncclConfig_t*) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

struct MgState {
    NcclApi api;
    ncclComm_t comm = nullptr;
    // A second communicator restricted to ONE CTA carries the pivots ahead of the panel.  A rank reaches the broadcast of
    // block b+1 at the START of its update with block b, milliseconds before the owner has factored that block, and an
    // NCCL kernel waits on the device: with the panel's own broadcast first in line, its CTAs (one per channel) sat
    // spinning on every non-owner rank under the trailing update and cost the DMMA GEMM ~10 % (per-rank timeline,
    // profiles/r2_mg_timeline.md).  Now the one-CTA broadcast does the waiting and the wide one starts when data flows.
    ncclComm_t comm_sig = nullptr;
    int rank = 0, nranks = 1;
    void* wbuf[2] = {nullptr, nullptr};  // packed panels (double buffered for the lookahead)
    size_t wbytes = 0;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_bcast[2] = {nullptr, nullptr};   // panel k landed in wbuf[k&1]
    cudaEvent_t ev_used[2] = {nullptr, nullptr};    // compute stream finished reading wbuf[k&1]
    cudaEvent_t ev_next = nullptr;                  // next block's columns are updated (owner only)
    cudaEvent_t ev_start = nullptr;
    // optional per-block-step timeline (lair_b200_mg_timeline): kTlPoints timing events per block step on the two streams
    bool timeline = false;
    std::vector<cudaEvent_t> tl;
    int64_t tl_nblk = 0;
};
// timeline points of block step b (all on this rank): what the main stream M and the panel / communication stream C were doing
enum TlPoint { kTlPanelReady = 0,   // M: the broadcast of block b has landed, the update with it may start
               kTlNextUpdated,      // M: (owner of b+1) block b+1's own columns are updated
               kTlPanelStart,       // C: start of factor + pack + broadcast of block b+1
               kTlPanelDone,        // C: (owner of b+1) block b+1 factored
               kTlPackDone,         // C: (owner of b+1) packed
               kTlBcastDone,        // C: broadcast of block b+1 complete on this rank
               kTlUpdateDone,       // M: trailing update with block b finished
               kTlStepDone,         // M: left interchanges done
               kTlPoints };
MgState g_mg;

int load_nccl(NcclApi& api) {
    if (api.handle) return LAIR_B200_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) {
        set_error("cannot load libnccl.so.2: %s", dlerror());
        return LAIR_B200_ERR_NCCL;
    }
#define LOAD(field, sym)                                                \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, sym)); \
    if (!api.field) {                                                   \
        set_error("libnccl is missing symbol %s", sym);                 \
        return LAIR_B200_ERR_NCCL;                                      \
    }
    LOAD(GetUniqueId, "ncclGetUniqueId")
    LOAD(CommInitRank, "ncclCommInitRank")
    LOAD(CommDestroy, "ncclCommDestroy")
    LOAD(CommSplit, "ncclCommSplit")
    LOAD(Broadcast, "ncclBroadcast")
    LOAD(AllReduce, "ncclAllReduce")
    LOAD(GroupStart, "ncclGroupStart")
    LOAD(GroupEnd, "ncclGroupEnd")
    LOAD(GetErrorString, "ncclGetErrorString")
#undef LOAD
    return LAIR_B200_OK;
}

#define LAIR_NCCL_CHECK(expr)                                                                        \
    do {                                                                                             \
        ncclResult_t _r = (expr);                                                                    \
        if (_r != ncclSuccess) {                                                                     \
            set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, g_mg.api.GetErrorString(_r));    \
            return LAIR_B200_ERR_NCCL;                                                               \
        }                                                                                            \
    } while (0)

// rows x w panel at src (leading dimension lds) -> packed rows x w at dst
template <class T>
__global__ void pack_panel_kernel(const T* __restrict__ src, long long lds, T* __restrict__ dst, long long rows, int w) {
    const long long total = rows * w;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / w;
        const int c = (int)(i - r * w);
        dst[i] = src[r * lds + c];
    }
}

// local-column bookkeeping of the block-cyclic distribution
struct Dist {
    int64_t n, nb;
    int rank, P;
    int64_t nblocks() const { return (n + nb - 1) / nb; }
    int owner(int64_t blk) const { return (int)(blk % P); }
    int64_t width(int64_t blk) const { return (n - blk * nb) < nb ? (n - blk * nb) : nb; }
    // local column index of the first local column whose global block index is >= blk
    int64_t first_local_col_of_block_at_or_after(int64_t blk) const {
        const int64_t cyc = blk / P, r = blk % P;
        int64_t lc = cyc * nb;
        if (rank < r) lc += nb;
        return lc < local_cols() ? lc : local_cols();
    }
    int64_t local_cols() const {
        int64_t c = 0;
        for (int64_t b = rank; b < nblocks(); b += P) c += width(b);
        return c;
    }
};

}  // namespace

template <class T>
int getrf_mg_dev(int64_t n, int64_t nb, T* d_a, int64_t lda, int32_t* d_ipiv, int32_t* d_info, cudaStream_t s) {
    MgState& mg = g_mg;
    LAIR_REQUIRE(mg.comm != nullptr, "getrf_mg: call lair_b200_mg_init first");
    LAIR_REQUIRE(n >= 1 && nb >= 32 && nb % 32 == 0, "getrf_mg: need n >= 1 and nb a positive multiple of 32");
    const Dist D{n, nb, mg.rank, mg.nranks};
    const int64_t lcols = D.local_cols();
    LAIR_REQUIRE(lda >= lcols, "getrf_mg: lda (%lld) smaller than the local column count (%lld)", (long long)lda, (long long)lcols);
    const ncclDataType_t dtype = sizeof(T) == 8 ? ncclFloat64 : ncclFloat32;

    // packed panel buffers
    const size_t need = (size_t)n * nb * sizeof(T);
    if (mg.wbytes < need) {
        for (auto& w : mg.wbuf) {
            if (w) LAIR_CUDA_CHECK(cudaFree(w));
            w = nullptr;
        }
        for (auto& w : mg.wbuf) LAIR_CUDA_CHECK(cudaMalloc(&w, need));
        mg.wbytes = need;
    }
    cudaStream_t M = s, C = mg.comm_stream;  // compute / (panel + communication, high priority)
    LAIR_CUDA_CHECK(cudaMemsetAsync(d_info, 0xFF, sizeof(int32_t), M));  // -1
    LAIR_CUDA_CHECK(cudaEventRecord(mg.ev_start, M));
    LAIR_CUDA_CHECK(cudaStreamWaitEvent(C, mg.ev_start, 0));

    const int64_t nblk = D.nblocks();
    if (mg.timeline) {
        const size_t need_ev = (size_t)(nblk + 1) * kTlPoints;
        while (mg.tl.size() < need_ev) {
            cudaEvent_t e;
            LAIR_CUDA_CHECK(cudaEventCreate(&e));
            mg.tl.push_back(e);
        }
        mg.tl_nblk = nblk;
        LAIR_CUDA_CHECK(cudaEventRecord(mg.tl[(size_t)nblk * kTlPoints], M));  // t = 0
    }
    auto mark = [&](int64_t blk, int point, cudaStream_t st) -> int {
        if (mg.timeline && blk >= 0 && blk < nblk) LAIR_CUDA_CHECK(cudaEventRecord(mg.tl[(size_t)blk * kTlPoints + point], st));
        return LAIR_B200_OK;
    };
    // factor + pack + broadcast of block `blk` on stream C (every rank calls this in the same order)
    auto panel_and_bcast = [&](int64_t blk) -> int {
        const int64_t j0 = blk * nb, w = D.width(blk), rows = n - j0;
        const int slot = (int)(blk & 1);
        T* wb = static_cast<T*>(mg.wbuf[slot]);
        // wbuf[slot] was last read by the compute stream during block blk-2
        if (blk >= 2) LAIR_CUDA_CHECK(cudaStreamWaitEvent(C, mg.ev_used[slot], 0));
        LAIR_CHECK(mark(blk - 1, kTlPanelStart, C));
        if (D.owner(blk) == D.rank) {
            const int64_t lc0 = (blk / D.P) * nb;
            LAIR_CHECK(getrf_block_dev<T>(n, d_a, lda, j0, lc0, w, d_ipiv, d_info, C));
            LAIR_CHECK(mark(blk - 1, kTlPanelDone, C));
            const long long total = rows * w;
            int grid = (int)((total + 255) / 256);
            if (grid > 148 * 8) grid = 148 * 8;
            pack_panel_kernel<T><<<grid, 256, 0, C>>>(d_a + j0 * lda + lc0, (long long)lda, wb, rows, (int)w);
            LAIR_LAUNCH_CHECK();
            LAIR_CHECK(mark(blk - 1, kTlPackDone, C));
        }
        if (mg.comm_sig) {
            LAIR_NCCL_CHECK(mg.api.Broadcast(d_ipiv + j0, d_ipiv + j0, (size_t)w, ncclInt32, D.owner(blk), mg.comm_sig, C));
            LAIR_NCCL_CHECK(mg.api.Broadcast(wb, wb, (size_t)rows * w, dtype, D.owner(blk), mg.comm, C));
        } else {
            LAIR_NCCL_CHECK(mg.api.GroupStart());
            LAIR_NCCL_CHECK(mg.api.Broadcast(wb, wb, (size_t)rows * w, dtype, D.owner(blk), mg.comm, C));
            LAIR_NCCL_CHECK(mg.api.Broadcast(d_ipiv + j0, d_ipiv + j0, (size_t)w, ncclInt32, D.owner(blk), mg.comm, C));
            LAIR_NCCL_CHECK(mg.api.GroupEnd());
        }
        LAIR_CUDA_CHECK(cudaEventRecord(mg.ev_bcast[slot], C));
        LAIR_CHECK(mark(blk - 1, kTlBcastDone, C));
        return LAIR_B200_OK;
    };
    // laswp + trsm + gemm of local columns [lc_a, lc_b) with the panel of block `blk` (in wbuf)
    auto update_local = [&](int64_t blk, int64_t lc_a, int64_t lc_b, cudaStream_t st) -> int {
        if (lc_b <= lc_a) return LAIR_B200_OK;
        const int64_t j0 = blk * nb, w = D.width(blk), r1 = j0 + w;
        const T* wb = static_cast<const T*>(mg.wbuf[blk & 1]);
        if (ctx().opt.trsm_strip != 0 && lc_b - lc_a > 512 && w <= 256 && (w > 64 || ctx().opt.trsm_strip == 2)) {
            // wide local ranges: interchanges as one bandwidth pass, the triangle in one register-tiled launch (trsm_strip.cu)
            LAIR_CHECK(laswp_dev<T>(lc_b - lc_a, d_a + lc_a, lda, j0, r1, d_ipiv, st));
            LAIR_CHECK(trsm_strip_dev<T>(w, lc_b - lc_a, wb, w, d_a + j0 * lda + lc_a, lda, st));
        } else if (ctx().opt.fuse_swap_trsm == 1 && w <= 256) {
            // the block step as a chain of 64-row fused laswp + prefix update + trsm launches (laswp_trsm.cu),
            // L taken from the packed panel: same arithmetic as the single-GPU sweep (blocked.cu)
            for (int64_t off = 0; off < w; off += 64) {
                const int64_t kk = (w - off) < 64 ? (w - off) : 64;
                LAIR_CHECK(laswp_trsm_dev<T>(lc_b - lc_a, d_a + lc_a, lda, j0 + off, kk, d_ipiv, wb + off * w + off, w, st, off));
            }
        } else {
            LAIR_CHECK(laswp_dev<T>(lc_b - lc_a, d_a + lc_a, lda, j0, r1, d_ipiv, st));
            LAIR_CHECK(trsm_lower_unit_dev<T>(w, lc_b - lc_a, wb, w, d_a + j0 * lda + lc_a, lda, st));
        }
        if (r1 < n)
            LAIR_CHECK(gemm_minus_dev<T>(n - r1, lc_b - lc_a, w, wb + w * w, w, d_a + j0 * lda + lc_a, lda, d_a + r1 * lda + lc_a, lda, st));
        return LAIR_B200_OK;
    };

    LAIR_CHECK(panel_and_bcast(0));
    for (int64_t blk = 0; blk < nblk; ++blk) {
        const int slot = (int)(blk & 1);
        const int64_t j0 = blk * nb, w = D.width(blk);
        LAIR_CUDA_CHECK(cudaStreamWaitEvent(M, mg.ev_bcast[slot], 0));
        LAIR_CHECK(mark(blk, kTlPanelReady, M));
        const int64_t lc_right = D.first_local_col_of_block_at_or_after(blk + 1);  // local columns right of block blk
        int64_t lc_after_next = lc_right;
        if (blk + 1 < nblk) {
            if (D.owner(blk + 1) == D.rank) {
                // lookahead: my block blk+1 first, so its panel path + broadcast overlap the rest
                lc_after_next = lc_right + D.width(blk + 1);
                LAIR_CHECK(update_local(blk, lc_right, lc_after_next, M));
                LAIR_CHECK(mark(blk, kTlNextUpdated, M));
                LAIR_CUDA_CHECK(cudaEventRecord(mg.ev_next, M));
                LAIR_CUDA_CHECK(cudaStreamWaitEvent(C, mg.ev_next, 0));
            }
            LAIR_CHECK(panel_and_bcast(blk + 1));
        }
        LAIR_CHECK(update_local(blk, lc_after_next, lcols, M));
        LAIR_CHECK(mark(blk, kTlUpdateDone, M));
        // interchanges reach back into the L part stored on this rank (columns of blocks < blk,
        // and -- on the owner -- nothing of block blk itself: the panel kernel already placed its rows)
        const int64_t lc_left_end = D.first_local_col_of_block_at_or_after(blk);
        if (lc_left_end > 0) LAIR_CHECK(laswp_dev<T>(lc_left_end, d_a, lda, j0, j0 + w, d_ipiv, M));
        LAIR_CHECK(mark(blk, kTlStepDone, M));
        LAIR_CUDA_CHECK(cudaEventRecord(mg.ev_used[slot], M));
    }
    // every rank ends with the same info: the last zero-pivot step seen by any panel owner
    LAIR_CUDA_CHECK(cudaEventRecord(mg.ev_next, M));
    LAIR_CUDA_CHECK(cudaStreamWaitEvent(C, mg.ev_next, 0));
    LAIR_NCCL_CHECK(mg.api.AllReduce(d_info, d_info, 1, ncclInt32, ncclMax, mg.comm, C));
    // the caller's stream also waits for the last communication-stream work
    LAIR_CUDA_CHECK(cudaEventRecord(mg.ev_next, C));
    LAIR_CUDA_CHECK(cudaStreamWaitEvent(M, mg.ev_next, 0));
    return LAIR_B200_OK;
}

}  // namespace lair

using namespace lair;

extern "C" {

static void reset_mg_state() { (void)lair_b200_mg_finalize(); }
static ResetHook g_mg_hook(reset_mg_state);

int lair_b200_mg_unique_id(void* id128) {
    LAIR_REQUIRE(id128 != nullptr, "mg_unique_id: null buffer");
    LAIR_CHECK(load_nccl(g_mg.api));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    ncclUniqueId id;
    LAIR_NCCL_CHECK(g_mg.api.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return LAIR_B200_OK;
}

int lair_b200_mg_init(int rank, int nranks, const void* id128) {
    LAIR_REQUIRE(id128 != nullptr && nranks >= 1 && rank >= 0 && rank < nranks, "mg_init: bad arguments");
    LAIR_CHECK(ensure_init());
    LAIR_CHECK(load_nccl(g_mg.api));
    MgState& mg = g_mg;
    if (mg.comm) {
        LAIR_REQUIRE(mg.rank == rank && mg.nranks == nranks, "mg_init: already initialised with a different rank/size");
        return LAIR_B200_OK;
    }
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    LAIR_NCCL_CHECK(mg.api.CommInitRank(&mg.comm, nranks, id, rank));
    mg.rank = rank;
    mg.nranks = nranks;
    if (nranks > 1 && ctx().opt.mg_signal_comm) {
        ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
        cfg.minCTAs = 1;
        cfg.maxCTAs = 1;
        LAIR_NCCL_CHECK(mg.api.CommSplit(mg.comm, 0, rank, &mg.comm_sig, &cfg));
    }
    int lo = 0, hi = 0;
    LAIR_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    LAIR_CUDA_CHECK(cudaStreamCreateWithPriority(&mg.comm_stream, cudaStreamNonBlocking, hi));
    for (auto& e : mg.ev_bcast) LAIR_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : mg.ev_used) LAIR_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    LAIR_CUDA_CHECK(cudaEventCreateWithFlags(&mg.ev_next, cudaEventDisableTiming));
    LAIR_CUDA_CHECK(cudaEventCreateWithFlags(&mg.ev_start, cudaEventDisableTiming));
    return LAIR_B200_OK;
}

int lair_b200_mg_finalize(void) {
    MgState& mg = g_mg;
    if (!mg.comm) return LAIR_B200_OK;
    cudaDeviceSynchronize();
    if (mg.comm_sig) mg.api.CommDestroy(mg.comm_sig);
    mg.comm_sig = nullptr;
    mg.api.CommDestroy(mg.comm);
    mg.comm = nullptr;
    for (auto& w : mg.wbuf) {
        if (w) cudaFree(w);
        w = nullptr;
    }
    mg.wbytes = 0;
    if (mg.comm_stream) cudaStreamDestroy(mg.comm_stream);
    mg.comm_stream = nullptr;
    for (auto& e : mg.ev_bcast) if (e) { cudaEventDestroy(e); e = nullptr; }
    for (auto& e : mg.ev_used) if (e) { cudaEventDestroy(e); e = nullptr; }
    if (mg.ev_next) { cudaEventDestroy(mg.ev_next); mg.ev_next = nullptr; }
    if (mg.ev_start) { cudaEventDestroy(mg.ev_start); mg.ev_start = nullptr; }
    for (auto& e : mg.tl) cudaEventDestroy(e);
    mg.tl.clear();
    mg.tl_nblk = 0;
    return LAIR_B200_OK;
}

int lair_b200_mg_timeline(int enable) {
    g_mg.timeline = enable != 0;
    return LAIR_B200_OK;
}

// out[b * 8 + p] = milliseconds from the start of the last getrf_mg_dev call to timeline point p of block step b on this
// rank (NaN: the point was not recorded on this rank, e.g. a non-owner's panel).  Returns the points through *nblk.
int lair_b200_mg_timeline_read(float* out, int64_t cap, int64_t* nblk_out) {
    MgState& mg = g_mg;
    LAIR_REQUIRE(out && nblk_out, "mg_timeline_read: null pointer");
    *nblk_out = mg.tl_nblk;
    LAIR_REQUIRE(cap >= mg.tl_nblk * kTlPoints, "mg_timeline_read: buffer too small");
    LAIR_CUDA_CHECK(cudaDeviceSynchronize());
    if (mg.tl_nblk == 0) return LAIR_B200_OK;
    cudaEvent_t t0 = mg.tl[(size_t)mg.tl_nblk * kTlPoints];
    for (int64_t i = 0; i < mg.tl_nblk * kTlPoints; ++i) {
        float ms = 0.f;
        cudaError_t e = cudaEventElapsedTime(&ms, t0, mg.tl[(size_t)i]);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            ms = __builtin_nanf("");
        }
        out[i] = ms;
    }
    return LAIR_B200_OK;
}

int lair_b200_dgetrf_mg_dev(int64_t n, int64_t nb, double* d_a_local, int64_t lda, int32_t* d_ipiv, int32_t* d_info, void* stream) {
    LAIR_CHECK(ensure_init());
    return getrf_mg_dev<double>(n, nb, d_a_local, lda, d_ipiv, d_info, (cudaStream_t)stream);
}

int lair_b200_sgetrf_mg_dev(int64_t n, int64_t nb, float* d_a_local, int64_t lda, int32_t* d_ipiv, int32_t* d_info, void* stream) {
    LAIR_CHECK(ensure_init());
    return getrf_mg_dev<float>(n, nb, d_a_local, lda, d_ipiv, d_info, (cudaStream_t)stream);
}

}  // extern "C"
