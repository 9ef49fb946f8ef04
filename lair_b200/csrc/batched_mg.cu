// Single-process multi-GPU entry of the batched LU ("batched small LU: independent units, contiguous slices of
// the batch, no collective"): lair_b200_{s,d}getrf_batched_mg(batch, n, a, ipiv, info, ngpu).
//
// The reference factors a batch by calling getrf per matrix (src/lapack/getrf.rs:46-120); a Rust caller holding one big
// host array can therefore use every GPU of the node from ONE process through this entry: device d of 0..ngpu-1 takes the
// contiguous slice batch_slice(batch, d, ngpu) and runs the same three-stream pipeline as the single-GPU entry (chunk
// i+1 travels up while chunk i is factored and chunk i-1 travels back), each over its own PCIe link.  The kernels are the
// single-GPU ones (batched_lu*.cu); nothing is exchanged between devices.
#include <mutex>

#include "common.cuh"

namespace lair {
namespace {

constexpr int kMaxDev = 16;
constexpr int kMaxChunksPerDev = 64;

struct BatchDev {
    bool ready = false;
    cudaStream_t up = nullptr, run = nullptr, down = nullptr;
    void* dA = nullptr;
    int32_t *dP = nullptr, *dI = nullptr;
    size_t capA = 0, capP = 0, capI = 0;
    cudaEvent_t landed[kMaxChunksPerDev] = {}, done[kMaxChunksPerDev] = {};
};
BatchDev g_bd[kMaxDev];

void reset_batch_devs() {
    int cur = 0;
    cudaGetDevice(&cur);
    for (int d = 0; d < kMaxDev; ++d) {
        BatchDev& b = g_bd[d];
        if (!b.ready) continue;
        cudaSetDevice(d);
        cudaDeviceSynchronize();
        if (b.dA) cudaFree(b.dA);
        if (b.dP) cudaFree(b.dP);
        if (b.dI) cudaFree(b.dI);
        for (auto& e : b.landed) if (e) cudaEventDestroy(e);
        for (auto& e : b.done) if (e) cudaEventDestroy(e);
        if (b.up) cudaStreamDestroy(b.up);
        if (b.run) cudaStreamDestroy(b.run);
        if (b.down) cudaStreamDestroy(b.down);
        b = BatchDev();
    }
    cudaSetDevice(cur);
}
ResetHook g_batch_mg_hook(reset_batch_devs);

int grow(void** p, size_t* cap, size_t need) {
    if (*cap >= need) return LAIR_B200_OK;
    if (*p) {
        LAIR_CUDA_CHECK(cudaDeviceSynchronize());
        LAIR_CUDA_CHECK(cudaFree(*p));
        *p = nullptr;
        *cap = 0;
    }
    const size_t want = need + (need >> 3) + 256;
    LAIR_CUDA_CHECK(cudaMalloc(p, want));
    *cap = want;
    return LAIR_B200_OK;
}

// queue the whole pipeline of device `d`'s slice; returns without waiting
template <class T>
int queue_slice(int d, int64_t batch, int64_t n, T* a, int32_t* ipiv, int32_t* info, int64_t per_chunk) {
    LAIR_CUDA_CHECK(cudaSetDevice(d));
    BatchDev& b = g_bd[d];
    if (!b.ready) {
        cudaDeviceProp prop;
        LAIR_CUDA_CHECK(cudaGetDeviceProperties(&prop, d));
        LAIR_REQUIRE(prop.major == 10, "getrf_batched_mg: device %d is sm_%d%d; this library is built for sm_100a (B200) only", d, prop.major, prop.minor);
        LAIR_CUDA_CHECK(cudaStreamCreateWithFlags(&b.up, cudaStreamNonBlocking));
        LAIR_CUDA_CHECK(cudaStreamCreateWithFlags(&b.run, cudaStreamNonBlocking));
        LAIR_CUDA_CHECK(cudaStreamCreateWithFlags(&b.down, cudaStreamNonBlocking));
        b.ready = true;
    }
    const size_t mat_bytes = (size_t)n * n * sizeof(T), piv_bytes = (size_t)n * sizeof(int32_t);
    LAIR_CHECK(grow(&b.dA, &b.capA, (size_t)batch * mat_bytes));
    LAIR_CHECK(grow(reinterpret_cast<void**>(&b.dP), &b.capP, (size_t)batch * piv_bytes));
    LAIR_CHECK(grow(reinterpret_cast<void**>(&b.dI), &b.capI, (size_t)batch * sizeof(int32_t)));
    int64_t per = per_chunk;
    if ((batch + per - 1) / per > kMaxChunksPerDev) per = (batch + kMaxChunksPerDev - 1) / kMaxChunksPerDev;
    const int nchunks = (int)((batch + per - 1) / per);
    for (int i = 0; i < nchunks; ++i) {
        if (!b.landed[i]) LAIR_CUDA_CHECK(cudaEventCreateWithFlags(&b.landed[i], cudaEventDisableTiming));
        if (!b.done[i]) LAIR_CUDA_CHECK(cudaEventCreateWithFlags(&b.done[i], cudaEventDisableTiming));
        const int64_t b0 = (int64_t)i * per, nb = (b0 + per <= batch) ? per : (batch - b0);
        char* dAi = (char*)b.dA + (size_t)b0 * mat_bytes;
        LAIR_CUDA_CHECK(cudaMemcpyAsync(dAi, (const char*)a + (size_t)b0 * mat_bytes, (size_t)nb * mat_bytes, cudaMemcpyHostToDevice, b.up));
        LAIR_CUDA_CHECK(cudaEventRecord(b.landed[i], b.up));
        LAIR_CUDA_CHECK(cudaStreamWaitEvent(b.run, b.landed[i], 0));
        LAIR_CHECK(getrf_batched_dev<T>(nb, n, (T*)dAi, b.dP + b0 * n, b.dI + b0, b.run));
        LAIR_CUDA_CHECK(cudaEventRecord(b.done[i], b.run));
        LAIR_CUDA_CHECK(cudaStreamWaitEvent(b.down, b.done[i], 0));
        LAIR_CUDA_CHECK(cudaMemcpyAsync((char*)a + (size_t)b0 * mat_bytes, dAi, (size_t)nb * mat_bytes, cudaMemcpyDeviceToHost, b.down));
    }
    (void)ipiv;
    (void)info;
    return LAIR_B200_OK;
}

// Pivots and info (n + 1 int32 per matrix) return in one piece behind the last chunk -- and only after EVERY device's chunk
// pipeline has been queued: callers usually hand in freshly allocated pageable arrays for them, and a pageable D2H blocks
// the issuing thread until the stream has drained, which would keep the next device from even starting.
template <class T>
int queue_results(int d, int64_t batch, int64_t n, int32_t* ipiv, int32_t* info) {
    LAIR_CUDA_CHECK(cudaSetDevice(d));
    BatchDev& b = g_bd[d];
    LAIR_CUDA_CHECK(cudaMemcpyAsync(ipiv, b.dP, (size_t)batch * n * sizeof(int32_t), cudaMemcpyDeviceToHost, b.down));
    LAIR_CUDA_CHECK(cudaMemcpyAsync(info, b.dI, (size_t)batch * sizeof(int32_t), cudaMemcpyDeviceToHost, b.down));
    return LAIR_B200_OK;
}

template <class T>
int getrf_batched_mg_host(int64_t batch, int64_t n, T* a, int32_t* ipiv, int32_t* info, int ngpu) {
    LAIR_REQUIRE(batch >= 0 && n >= 0 && n <= 32, "getrf_batched_mg: need batch >= 0 and 0 <= n <= 32");
    LAIR_REQUIRE(ngpu >= 1 && ngpu <= kMaxDev, "getrf_batched_mg: ngpu must be in 1..%d, got %d", kMaxDev, ngpu);
    if (batch == 0 || n == 0) return LAIR_B200_OK;
    LAIR_REQUIRE(a && ipiv && info, "getrf_batched_mg: null pointer");
    std::lock_guard<std::mutex> lk(host_call_mutex());
    LAIR_CHECK(ensure_init());
    int count = 0;
    LAIR_CUDA_CHECK(cudaGetDeviceCount(&count));
    LAIR_REQUIRE(ngpu <= count, "getrf_batched_mg: %d GPUs requested, %d visible", ngpu, count);
    const int home = ctx().device;
    const int64_t per_chunk = ctx().opt.batched_chunk;
    int status = LAIR_B200_OK;
    // contiguous slices, shares differ by at most one matrix (lair_b200/sharding.py: batch_slice)
    const int64_t base = batch / ngpu, rem = batch % ngpu;
    int used = 0;
    for (int d = 0; d < ngpu && status == LAIR_B200_OK; ++d) {
        const int64_t start = d * base + (d < rem ? d : rem), cnt = base + (d < rem ? 1 : 0);
        if (cnt == 0) continue;
        status = queue_slice<T>(d, cnt, n, a + (size_t)start * n * n, ipiv + start * n, info + start, per_chunk);
        used = d + 1;
    }
    for (int d = 0; d < used && status == LAIR_B200_OK; ++d) {
        const int64_t start = d * base + (d < rem ? d : rem), cnt = base + (d < rem ? 1 : 0);
        if (cnt > 0) status = queue_results<T>(d, cnt, n, ipiv + start * n, info + start);
    }
    // wait for every device that was given work, also on failure (nothing may still write into the caller's arrays)
    for (int d = 0; d < used; ++d) {
        if (!g_bd[d].ready) continue;
        cudaSetDevice(d);
        const cudaError_t e = cudaStreamSynchronize(g_bd[d].down);
        const cudaError_t e2 = cudaStreamSynchronize(g_bd[d].run);
        if (status == LAIR_B200_OK && (e != cudaSuccess || e2 != cudaSuccess)) {
            set_error("getrf_batched_mg: device %d failed: %s", d, cudaGetErrorString(e != cudaSuccess ? e : e2));
            status = LAIR_B200_ERR_CUDA;
        }
    }
    cudaSetDevice(home);
    return status;
}

}  // namespace
}  // namespace lair

using namespace lair;

extern "C" {

int lair_b200_sgetrf_batched_mg(int64_t batch, int64_t n, float* a, int32_t* ipiv, int32_t* info, int ngpu) {
    return getrf_batched_mg_host<float>(batch, n, a, ipiv, info, ngpu);
}
int lair_b200_dgetrf_batched_mg(int64_t batch, int64_t n, double* a, int32_t* ipiv, int32_t* info, int ngpu) {
    return getrf_batched_mg_host<double>(batch, n, a, ipiv, info, ngpu);
}

}  // extern "C"
