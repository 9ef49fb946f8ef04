// Optional per-kernel-family timing with CUDA events on the launching stream, used by
// bench.py to report the dominant kernel's achieved FLOP/s (or GB/s) live.  Off by default;
// when off a ProfScope costs one branch.
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace lair {

struct ProfRecord {
    int bucket;
    cudaEvent_t e0, e1;
    double work;
};

static bool g_prof_on = false;
static std::vector<ProfRecord> g_prof_records;
static std::vector<cudaEvent_t> g_prof_pool;
static std::mutex g_prof_mu;
static double g_prof_ms[kProfBuckets];
static double g_prof_work[kProfBuckets];
static int64_t g_prof_count[kProfBuckets];

static const char* kBucketNames[kProfBuckets] = {"gemm", "panel", "laswp", "trsm", "batched", "small", "other"};

bool prof_enabled() { return g_prof_on; }

static cudaEvent_t prof_event() {
    if (!g_prof_pool.empty()) {
        cudaEvent_t e = g_prof_pool.back();
        g_prof_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

ProfScope::ProfScope(int bucket, cudaStream_t s, double work) : bucket_(bucket), stream_(s), work_(work), e0_(nullptr) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    e0_ = prof_event();
    cudaEventRecord(e0_, stream_);
}

ProfScope::~ProfScope() {
    if (!e0_) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEvent_t e1 = prof_event();
    cudaEventRecord(e1, stream_);
    g_prof_records.push_back({bucket_, e0_, e1, work_});
}

}  // namespace lair

using namespace lair;

extern "C" {

int lair_b200_profile_begin(void) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_prof_records) {
        g_prof_pool.push_back(r.e0);
        g_prof_pool.push_back(r.e1);
    }
    g_prof_records.clear();
    memset(g_prof_ms, 0, sizeof(g_prof_ms));
    memset(g_prof_work, 0, sizeof(g_prof_work));
    memset(g_prof_count, 0, sizeof(g_prof_count));
    g_prof_on = true;
    return LAIR_B200_OK;
}

int lair_b200_profile_end(void) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on = false;
    LAIR_CUDA_CHECK(cudaDeviceSynchronize());
    for (auto& r : g_prof_records) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
            g_prof_ms[r.bucket] += ms;
            g_prof_work[r.bucket] += r.work;
            g_prof_count[r.bucket] += 1;
        }
        g_prof_pool.push_back(r.e0);
        g_prof_pool.push_back(r.e1);
    }
    g_prof_records.clear();
    return LAIR_B200_OK;
}

int lair_b200_profile_get(const char* bucket, double* ms, int64_t* launches, double* work) {
    if (!bucket) return LAIR_B200_ERR_INVALID;
    for (int b = 0; b < kProfBuckets; ++b) {
        if (!strcmp(bucket, kBucketNames[b])) {
            if (ms) *ms = g_prof_ms[b];
            if (launches) *launches = g_prof_count[b];
            if (work) *work = g_prof_work[b];
            return LAIR_B200_OK;
        }
    }
    set_error("unknown profile bucket '%s'", bucket);
    return LAIR_B200_ERR_INVALID;
}

}  // extern "C"
