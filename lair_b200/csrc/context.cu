// Process-wide context: device binding, streams, workspaces, options, error string.
#include <atomic>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"
#include "host_io.cuh"

namespace lair {

thread_local int g_last_status = 0;
static thread_local char g_err[1024] = {0};
static std::atomic<int64_t> g_launches{0};
static Context g_ctx;
static std::mutex g_ctx_mu;
static std::atomic<uint64_t> g_epoch{1};
static constexpr int kMaxHooks = 16;
static void (*g_hooks[kMaxHooks])() = {nullptr};
static std::atomic<int> g_nhooks{0};

uint64_t context_epoch() { return g_epoch.load(std::memory_order_relaxed); }
void register_reset_hook(void (*fn)()) {
    const int i = g_nhooks.fetch_add(1);
    if (i < kMaxHooks) g_hooks[i] = fn;
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

Context& ctx() { return g_ctx; }

static int init_locked(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error("no CUDA device available (%s); lair_b200 has no CPU fallback",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        (void)cudaGetLastError();
        return LAIR_B200_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= count) {
        set_error("device %d out of range (0..%d)", device, count - 1);
        return LAIR_B200_ERR_INVALID;
    }
    if (g_ctx.ready && g_ctx.device == device) {
        LAIR_CUDA_CHECK(cudaSetDevice(device));
        return LAIR_B200_OK;
    }
    if (g_ctx.ready) {
        set_error("context already bound to device %d; call lair_b200_shutdown() first", g_ctx.device);
        return LAIR_B200_ERR_INVALID;
    }
    LAIR_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    LAIR_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major,
                  prop.minor);
        return LAIR_B200_ERR_NO_DEVICE;
    }
    g_ctx.device = device;
    g_ctx.sm_count = prop.multiProcessorCount;
    g_ctx.cc_major = prop.major;
    g_ctx.cc_minor = prop.minor;
    g_ctx.smem_optin = prop.sharedMemPerBlockOptin;
    int lo = 0, hi = 0;
    LAIR_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    LAIR_CUDA_CHECK(cudaStreamCreateWithPriority(&g_ctx.stream, cudaStreamNonBlocking, lo));
    // the lookahead panel runs on the high-priority stream so its CTAs are placed first
    LAIR_CUDA_CHECK(cudaStreamCreateWithPriority(&g_ctx.aux_stream, cudaStreamNonBlocking, hi));
    for (auto& ev : g_ctx.ev) LAIR_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    LAIR_CUDA_CHECK(cudaStreamCreateWithFlags(&g_ctx.copy_stream, cudaStreamNonBlocking));
    LAIR_CUDA_CHECK(cudaStreamCreateWithFlags(&g_ctx.drain_stream, cudaStreamNonBlocking));
    LAIR_CUDA_CHECK(cudaMalloc(&g_ctx.d_fault, 64));
    LAIR_CUDA_CHECK(cudaMemset(g_ctx.d_fault, 0, 64));
    if (const char* v = getenv("LAIR_B200_NB")) g_ctx.opt.nb = atoll(v);
    if (const char* v = getenv("LAIR_B200_SMALL_N")) g_ctx.opt.small_n = atoll(v);
    if (const char* v = getenv("LAIR_B200_LOOKAHEAD")) g_ctx.opt.lookahead = atoll(v);
    if (const char* v = getenv("LAIR_B200_BATCHED_CFG")) g_ctx.opt.batched_cfg = atoll(v);
    if (const char* v = getenv("LAIR_B200_MG_SIGNAL_COMM")) g_ctx.opt.mg_signal_comm = atoll(v);
    g_epoch.fetch_add(1);  // per-device caches in function-local statics are stale from here on
    g_ctx.ready = true;
    return LAIR_B200_OK;
}

int ensure_init() {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    if (g_ctx.ready) {
        LAIR_CUDA_CHECK(cudaSetDevice(g_ctx.device));
        return LAIR_B200_OK;
    }
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        dev = 0;
    }
    return init_locked(dev);
}

int check_fault(cudaStream_t s) {
    Context& c = g_ctx;
    if (!c.ready || !c.d_fault) return LAIR_B200_OK;
    int v = 0;
    LAIR_CUDA_CHECK(cudaMemcpyAsync(&v, c.d_fault, sizeof(int), cudaMemcpyDeviceToHost, s));
    LAIR_CUDA_CHECK(cudaStreamSynchronize(s));
    if (v == 0) return LAIR_B200_OK;
    LAIR_CUDA_CHECK(cudaMemsetAsync(c.d_fault, 0, sizeof(int), s));
    LAIR_CUDA_CHECK(cudaStreamSynchronize(s));
    set_error("a cross-CTA wait timed out on the device (fault %d: %s); the results of this call are invalid", v,
              v == 1 ? "tall-panel pivot exchange" : "dataflow triangular solve");
    return LAIR_B200_ERR_CUDA;
}

int ensure_scratch(size_t bytes, void** out) {
    Context& c = g_ctx;
    if (c.scratch_bytes < bytes) {
        if (c.scratch) {
            LAIR_CUDA_CHECK(cudaDeviceSynchronize());
            LAIR_CUDA_CHECK(cudaFree(c.scratch));
            c.scratch = nullptr;
            c.scratch_bytes = 0;
        }
        size_t want = bytes + (bytes >> 2);
        LAIR_CUDA_CHECK(cudaMalloc(&c.scratch, want));
        c.scratch_bytes = want;
    }
    *out = c.scratch;
    return LAIR_B200_OK;
}

int ensure_work(int slot, size_t bytes, void** out, cudaStream_t s) {
    Context& c = g_ctx;
    LAIR_REQUIRE(slot >= 0 && slot < Context::kWorkSlots, "ensure_work: bad slot %d", slot);
    if (c.work_bytes[slot] < bytes) {
        if (c.work[slot]) {
            LAIR_CUDA_CHECK(cudaStreamSynchronize(s));
            LAIR_CUDA_CHECK(cudaFree(c.work[slot]));
            c.work[slot] = nullptr;
            c.work_bytes[slot] = 0;
        }
        const size_t want = bytes + (bytes >> 2) + 256;
        LAIR_CUDA_CHECK(cudaMalloc(&c.work[slot], want));
        c.work_bytes[slot] = want;
    }
    *out = c.work[slot];
    return LAIR_B200_OK;
}

}  // namespace lair

using namespace lair;

extern "C" {

int lair_b200_version(void) { return 100; }  // 0.1.0

const char* lair_b200_last_error(void) { return g_err; }

int lair_b200_device_count(int* count) {
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        c = 0;
    }
    if (count) *count = c;
    if (c == 0) {
        set_error("no CUDA device available; lair_b200 has no CPU fallback");
        return LAIR_B200_ERR_NO_DEVICE;
    }
    return LAIR_B200_OK;
}

int lair_b200_init(int device) {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    return init_locked(device);
}

int lair_b200_shutdown(void) {
    // lock order as in the host-pointer entry points (call lock, then context lock): none of them is in flight while state goes away
    std::lock_guard<std::mutex> call_lk(host_call_mutex());
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    Context& c = g_ctx;
    if (!c.ready) return LAIR_B200_OK;
    cudaSetDevice(c.device);
    cudaDeviceSynchronize();
    for (int i = 0; i < g_nhooks.load() && i < kMaxHooks; ++i)
        if (g_hooks[i]) g_hooks[i]();  // module-owned device buffers (laswp_perm, trsm_ll, mg) are freed and forgotten
    if (c.panel_ws) cudaFree(c.panel_ws);
    if (c.scratch) cudaFree(c.scratch);
    for (auto& w : c.work)
        if (w) cudaFree(w);
    for (auto& ev : c.ev)
        if (ev) cudaEventDestroy(ev);
    if (c.stream) cudaStreamDestroy(c.stream);
    if (c.aux_stream) cudaStreamDestroy(c.aux_stream);
    if (c.copy_stream) cudaStreamDestroy(c.copy_stream);
    if (c.drain_stream) cudaStreamDestroy(c.drain_stream);
    for (auto& ev : c.drain_ev)
        if (ev) cudaEventDestroy(ev);
    if (c.d_fault) cudaFree(c.d_fault);
    for (auto& ev : c.chunk_ev)
        if (ev) cudaEventDestroy(ev);
    pool().release();
    c = Context();
    return LAIR_B200_OK;
}

int lair_b200_set_option(const char* name, int64_t value) {
    if (!name) return LAIR_B200_ERR_INVALID;
    Options& o = g_ctx.opt;
    if (!strcmp(name, "nb")) {
        if (value != 0 && (value < 32 || value % 32)) {
            set_error("nb must be 0 (automatic) or a positive multiple of 32, got %lld", (long long)value);
            return LAIR_B200_ERR_INVALID;
        }
        o.nb = value;
    } else if (!strcmp(name, "nb_t1")) {
        o.nb_t1 = value;
    } else if (!strcmp(name, "nb_t2")) {
        o.nb_t2 = value;
    } else if (!strcmp(name, "small_n")) {
        o.small_n = value;
    } else if (!strcmp(name, "debug_raise_fault")) {
        // test hook: raise the device fault word exactly as a timed-out wait would (tests/test_gpu_parity.py)
        LAIR_CHECK(ensure_init());
        const int v = (int)value;
        LAIR_CUDA_CHECK(cudaMemcpy(g_ctx.d_fault, &v, sizeof(int), cudaMemcpyHostToDevice));
    } else if (!strcmp(name, "lookahead")) {
        o.lookahead = value;
    } else if (!strcmp(name, "chain_on_p")) {
        o.chain_on_p = value;
    } else if (!strcmp(name, "batched_cfg")) {
        o.batched_cfg = value;
    } else if (!strcmp(name, "panel_cluster")) {
        o.panel_cluster = value;
    } else if (!strcmp(name, "qr_blocked")) {
        o.qr_blocked = value;
    } else if (!strcmp(name, "cx_blocked")) {
        o.cx_blocked = value;
    } else if (!strcmp(name, "gemm_cfg")) {
        o.gemm_cfg = value;
    } else if (!strcmp(name, "pair_small")) {
        o.pair_small = value;
    } else if (!strcmp(name, "pair_small_f32")) {
        o.pair_small_f32 = value;
    } else if (!strcmp(name, "pair_k512")) {
        o.pair_k512 = value;
    } else if (!strcmp(name, "trsm_strip")) {
        o.trsm_strip = value;
    } else if (!strcmp(name, "drain_rows")) {
        o.drain_rows = value;
    } else if (!strcmp(name, "mg_signal_comm")) {
        o.mg_signal_comm = value;
    } else if (!strcmp(name, "sgemm_tf32")) {
        o.sgemm_tf32 = value;
    } else if (!strcmp(name, "gemm_raster")) {
        o.gemm_raster = value < 1 ? 1 : value;
    } else if (!strcmp(name, "panel_group")) {
        o.panel_group = value;
    } else if (!strcmp(name, "panel_rpt")) {
        o.panel_rpt = value;
    } else if (!strcmp(name, "panel_timing")) {
        o.panel_timing = value;
    } else if (!strcmp(name, "trsm_dataflow")) {
        o.trsm_dataflow = value;
    } else if (!strcmp(name, "stream_cols")) {
        if (value != 0 && (value < 256 || value % 256)) {
            set_error("stream_cols must be 0 or a positive multiple of 256, got %lld", (long long)value);
            return LAIR_B200_ERR_INVALID;
        }
        o.stream_cols = value;
    } else if (!strcmp(name, "batched_chunk")) {
        if (value < 256) {
            set_error("batched_chunk must be >= 256 matrices, got %lld", (long long)value);
            return LAIR_B200_ERR_INVALID;
        }
        o.batched_chunk = value;
    } else if (!strcmp(name, "stream_join_div")) {
        o.stream_join_div = value < 1 ? 1 : value;
    } else if (!strcmp(name, "laswp_perm")) {
        o.laswp_perm = value;
    } else if (!strcmp(name, "trsm_rb")) {
        o.trsm_rb = value;
    } else if (!strcmp(name, "fuse_swap_trsm")) {
        o.fuse_swap_trsm = value;
    } else if (!strcmp(name, "panel_exchange")) {
        o.panel_exchange = value;
    } else if (!strcmp(name, "panel_w64")) {
        o.panel_w64 = value;
    } else {
        set_error("unknown option '%s'", name);
        return LAIR_B200_ERR_INVALID;
    }
    return LAIR_B200_OK;
}

int lair_b200_get_option(const char* name, int64_t* value) {
    if (!name || !value) return LAIR_B200_ERR_INVALID;
    const Options& o = g_ctx.opt;
    if (!strcmp(name, "nb")) *value = o.nb;
    else if (!strcmp(name, "nb_t1")) *value = o.nb_t1;
    else if (!strcmp(name, "nb_t2")) *value = o.nb_t2;
    else if (!strcmp(name, "small_n")) *value = o.small_n;
    else if (!strcmp(name, "lookahead")) *value = o.lookahead;
    else if (!strcmp(name, "chain_on_p")) *value = o.chain_on_p;
    else if (!strcmp(name, "batched_cfg")) *value = o.batched_cfg;
    else if (!strcmp(name, "panel_cluster")) *value = o.panel_cluster;
    else if (!strcmp(name, "qr_blocked")) *value = o.qr_blocked;
    else if (!strcmp(name, "cx_blocked")) *value = o.cx_blocked;
    else if (!strcmp(name, "gemm_cfg")) *value = o.gemm_cfg;
    else if (!strcmp(name, "gemm_raster")) *value = o.gemm_raster;
    else if (!strcmp(name, "sgemm_tf32")) *value = o.sgemm_tf32;
    else if (!strcmp(name, "pair_small")) *value = o.pair_small;
    else if (!strcmp(name, "pair_small_f32")) *value = o.pair_small_f32;
    else if (!strcmp(name, "pair_k512")) *value = o.pair_k512;
    else if (!strcmp(name, "trsm_strip")) *value = o.trsm_strip;
    else if (!strcmp(name, "drain_rows")) *value = o.drain_rows;
    else if (!strcmp(name, "mg_signal_comm")) *value = o.mg_signal_comm;
    else if (!strcmp(name, "panel_group")) *value = o.panel_group;
    else if (!strcmp(name, "panel_rpt")) *value = o.panel_rpt;
    else if (!strcmp(name, "panel_timing")) *value = o.panel_timing;
    else if (!strcmp(name, "trsm_dataflow")) *value = o.trsm_dataflow;
    else if (!strcmp(name, "stream_cols")) *value = o.stream_cols;
    else if (!strcmp(name, "batched_chunk")) *value = o.batched_chunk;
    else if (!strcmp(name, "stream_join_div")) *value = o.stream_join_div;
    else if (!strcmp(name, "laswp_perm")) *value = o.laswp_perm;
    else if (!strcmp(name, "trsm_rb")) *value = o.trsm_rb;
    else if (!strcmp(name, "fuse_swap_trsm")) *value = o.fuse_swap_trsm;
    else if (!strcmp(name, "panel_exchange")) *value = o.panel_exchange;
    else if (!strcmp(name, "panel_w64")) *value = o.panel_w64;
    else {
        set_error("unknown option '%s'", name);
        return LAIR_B200_ERR_INVALID;
    }
    return LAIR_B200_OK;
}

int64_t lair_b200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int lair_b200_check_fault(void* stream) { return check_fault(static_cast<cudaStream_t>(stream)); }

int lair_b200_debug_panel_timing(long long* out8, int clear) {
    if (g_ctx.opt.panel_cluster >= 3) return panel_push_timing(out8, clear != 0);
    if (g_ctx.opt.panel_cluster >= 2) return panel_blocked_timing(out8, clear != 0);
    return panel_cluster_timing(out8, clear != 0);
}

}  // extern "C"
