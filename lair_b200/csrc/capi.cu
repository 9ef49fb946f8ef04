// extern "C" entry points declared in include/lair_b200.h: argument checks, dispatch between
// the single-CTA exact kernel and the blocked factorization, host <-> device marshalling.
#include <mutex>
#include <new>
#include <vector>

#include "common.cuh"
#include "host_io.cuh"

namespace lair {

static std::mutex g_call_mu;  // host-pointer entry points serialise on the library stream
std::mutex& host_call_mutex() { return g_call_mu; }

// ---- dispatch ------------------------------------------------------------------------------
template <class T> struct IsReal { static constexpr bool value = !Ops<T>::is_complex; };

template <class T>
static bool fits_small(int64_t m, int64_t n) {
    const int64_t lim = ctx().opt.small_n;
    return (m <= lim && n <= lim);
}

template <class T>
int getrf_dev(int64_t m, int64_t n, T* d_a, int64_t lda, int32_t* d_ipiv, int32_t* d_info, bool std_layout,
              cudaStream_t s, const ColumnFeed* feed, RowDrain* drain) {
    if (fits_small<T>(m, n)) return getrf_small_dev<T>(m, n, d_a, lda, d_ipiv, d_info, std_layout, s);
    if constexpr (IsReal<T>::value) {
        return getrf_blocked_dev<T>(m, n, d_a, lda, d_ipiv, d_info, s, feed, drain);
    } else {
        // complex beyond the single-CTA limit: blocked sweep, ZGEMM as one real GEMM on packed operands (blocked_cx.cu)
        if (ctx().opt.cx_blocked == 0) return getrf_small_dev<T>(m, n, d_a, lda, d_ipiv, d_info, std_layout, s);
        return getrf_blocked_cx_dev<T>(m, n, d_a, lda, d_ipiv, d_info, std_layout, s);
    }
}

template <class T>
int getrs_dev(int64_t n, int64_t nrhs, const T* d_lu, int64_t lda, const int32_t* d_ipiv, T* d_b, int64_t ldb,
              cudaStream_t s) {
    if (n <= ctx().opt.small_n) return getrs_small_dev<T>(n, nrhs, d_lu, lda, d_ipiv, d_b, ldb, s);
    if constexpr (IsReal<T>::value) {
        return getrs_blocked_dev<T>(n, nrhs, d_lu, lda, d_ipiv, d_b, ldb, s);
    } else {
        if (ctx().opt.cx_blocked == 0) return getrs_small_dev<T>(n, nrhs, d_lu, lda, d_ipiv, d_b, ldb, s);
        return getrs_blocked_cx_dev<T>(n, nrhs, d_lu, lda, d_ipiv, d_b, ldb, s);
    }
}

static int64_t device_ld(int64_t n) {
    // keep rows 128-byte friendly for the blocked kernels; small matrices stay dense
    if (n <= ctx().opt.small_n) return n;
    return (n + 31) / 32 * 32;
}

// ---- chunked upload: the factorization starts when the first column chunk has landed ---------
// Returns true (and fills `feed`) when the matrix was queued on the copy stream in column chunks;
// false when the caller should upload it in one piece (small, complex, or not row-contiguous).
template <class T>
static int upload_matrix_chunked(const T* a, int64_t m, int64_t n, int64_t rs, int64_t cs, T* d, int64_t ld, ColumnFeed* feed,
                                 bool* chunked) {
    *chunked = false;
    const int64_t w = ctx().opt.stream_cols;
    // the sweep assumes the first chunk holds the first two blocks (blocked.cu: join rule cs < c0 + 2 nb + 512): a forced block
    // width above 512 or above half a chunk would touch columns whose copy has not been waited for -- upload in one piece then
    const int64_t nb_forced = ctx().opt.nb;
    if (nb_forced > 512 || 2 * nb_forced > w) return LAIR_B200_OK;
    if (!IsReal<T>::value || w < 512 || fits_small<T>(m, n) || cs != 1 || rs < n || n < 2 * w || m < n / 2) return LAIR_B200_OK;
    const int nchunks = (int)((n + w - 1) / w);
    if (nchunks > Context::kMaxChunks) return LAIR_B200_OK;
    Context& c = ctx();
    for (int i = 0; i < nchunks; ++i) {
        if (!c.chunk_ev[i]) LAIR_CUDA_CHECK(cudaEventCreateWithFlags(&c.chunk_ev[i], cudaEventDisableTiming));
        const int64_t c0 = (int64_t)i * w, cw = (c0 + w <= n) ? w : (n - c0);
        LAIR_CUDA_CHECK(cudaMemcpy2DAsync(d + c0, (size_t)ld * sizeof(T), a + c0, (size_t)rs * sizeof(T), (size_t)cw * sizeof(T),
                                          (size_t)m, cudaMemcpyHostToDevice, c.copy_stream));
        LAIR_CUDA_CHECK(cudaEventRecord(c.chunk_ev[i], c.copy_stream));
    }
    feed->chunk = w;
    feed->nchunks = nchunks;
    feed->ready = c.chunk_ev;
    *chunked = true;
    return LAIR_B200_OK;
}

// ---- host-pointer getrf --------------------------------------------------------------------
template <class T>
static int getrf_host(int64_t m, int64_t n, T* a, int64_t rs, int64_t cs, int64_t* ipiv, int64_t* info) {
    LAIR_REQUIRE(m >= 0 && n >= 0, "getrf: negative dimension (m=%lld, n=%lld)", (long long)m, (long long)n);
    LAIR_REQUIRE(info != nullptr, "getrf: info is null");
    const int64_t k = m < n ? m : n;
    *info = -1;
    if (k == 0) return LAIR_B200_OK;  // getrf.rs:350-355: empty -> ([], None)
    LAIR_REQUIRE(a != nullptr && ipiv != nullptr, "getrf: null pointer");
    std::lock_guard<std::mutex> lk(g_call_mu);
    LAIR_CHECK(ensure_init());
    cudaStream_t s = ctx().stream;
    const bool std_layout = is_standard_layout(m, n, rs, cs);
    const int64_t ld = device_ld(n);
    void *dA = nullptr, *dP = nullptr, *dI = nullptr;
    LAIR_CHECK(pool().get(DevicePool::kMatrix, (size_t)m * ld * sizeof(T), &dA));
    LAIR_CHECK(pool().get(DevicePool::kPivots, (size_t)k * sizeof(int32_t), &dP));
    LAIR_CHECK(pool().get(DevicePool::kInfo, sizeof(int32_t), &dI));
    ColumnFeed feed;
    bool chunked = false;
    LAIR_CHECK(upload_matrix_chunked<T>(a, m, n, rs, cs, (T*)dA, ld, &feed, &chunked));
    if (!chunked) LAIR_CHECK(upload_matrix<T>(a, m, n, rs, cs, (T*)dA, ld, DevicePool::kTmpA, s));
    // finished rows go home while the sweep still runs: real types, row-contiguous PINNED host array (a pageable target would
    // block this thread inside every copy), large enough for the copies to matter
    RowDrain drain;
    bool draining = false;
    if (IsReal<T>::value && ctx().opt.drain_rows != 0 && !fits_small<T>(m, n) && cs == 1 && rs >= n && m >= 2048 && n >= 2048) {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, a) == cudaSuccess && pa.type == cudaMemoryTypeHost) {
            drain.host = a;
            drain.host_rs = rs;
            drain.stream = ctx().drain_stream;
            drain.ev = ctx().drain_ev;
            drain.nev = Context::kMaxChunks - 1;  // the last event is the caller's (below)
            const int64_t by_bytes = ((int64_t)8 << 20) / (n * (int64_t)sizeof(T)) + 1, by_count = (m + Context::kMaxChunks - 2) / (Context::kMaxChunks - 1);
            drain.min_rows = by_bytes > by_count ? by_bytes : by_count;
            draining = true;
        } else {
            (void)cudaGetLastError();
        }
    }
    LAIR_CHECK(getrf_dev<T>(m, n, (T*)dA, ld, (int32_t*)dP, (int32_t*)dI, std_layout, s, chunked ? &feed : nullptr, draining ? &drain : nullptr));
    if (draining && drain.drained > 0) {
        const int64_t r0 = drain.drained;  // the rows the sweep finished last
        if (r0 < m)
            LAIR_CUDA_CHECK(cudaMemcpy2DAsync(a + r0 * rs, (size_t)rs * sizeof(T), (const T*)dA + r0 * ld, (size_t)ld * sizeof(T), (size_t)n * sizeof(T),
                                              (size_t)(m - r0), cudaMemcpyDeviceToHost, s));
        cudaEvent_t& ev = ctx().drain_ev[Context::kMaxChunks - 1];  // (never handed to the sweep: nev events start at index 0 and min_rows caps their number)
        if (!ev) LAIR_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        LAIR_CUDA_CHECK(cudaEventRecord(ev, drain.stream));
        LAIR_CUDA_CHECK(cudaStreamWaitEvent(s, ev, 0));  // the wait on `s` below covers the drained rows too
    } else {
        LAIR_CHECK(download_matrix<T>(a, m, n, rs, cs, (const T*)dA, ld, DevicePool::kTmpA, s));
    }
    LAIR_CHECK(download_ipiv64(ipiv, (const int32_t*)dP, k, DevicePool::kPivots64, s));
    int32_t info32 = -1;
    LAIR_CUDA_CHECK(cudaMemcpyAsync(&info32, dI, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    LAIR_CHECK(check_fault(s));  // waits for the stream; fails loudly if a device-side wait timed out
    *info = info32;
    return LAIR_B200_OK;
}

// ---- host-pointer getrs / gesv --------------------------------------------------------------
template <class T>
static int upload_ipiv32(const int64_t* ipiv, int64_t n, int32_t* d, cudaStream_t s) {
    std::vector<int32_t> p((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        // LAPACK-style interchanges (what getrf returns): step i exchanges row i with a row at or below it.  The device
        // kernels track only rows >= the current step, so an in-range entry ABOVE its step is refused, not mis-applied.
        LAIR_REQUIRE(ipiv[i] >= i && ipiv[i] < n, "getrs: ipiv[%lld]=%lld is not in [%lld, %lld): the pivots must be getrf's sequential interchanges",
                     (long long)i, (long long)ipiv[i], (long long)i, (long long)n);
        p[(size_t)i] = (int32_t)ipiv[i];
    }
    LAIR_CUDA_CHECK(cudaMemcpyAsync(d, p.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    LAIR_CUDA_CHECK(cudaStreamSynchronize(s));
    return LAIR_B200_OK;
}

template <class T>
static int getrs_host(int64_t n, int64_t nrhs, const T* lu, int64_t lu_rs, int64_t lu_cs, const int64_t* ipiv, const T* b,
                      int64_t b_rs, int64_t b_cs, T* x, int64_t x_rs, int64_t x_cs) {
    LAIR_REQUIRE(n >= 0 && nrhs >= 0, "getrs: negative dimension");
    if (n == 0 || nrhs == 0) return LAIR_B200_OK;
    LAIR_REQUIRE(lu && ipiv && b && x, "getrs: null pointer");
    std::lock_guard<std::mutex> lk(g_call_mu);
    LAIR_CHECK(ensure_init());
    cudaStream_t s = ctx().stream;
    const int64_t ld = device_ld(n);
    const int64_t ldb = nrhs;
    void *dA = nullptr, *dB = nullptr, *dP = nullptr;
    LAIR_CHECK(pool().get(DevicePool::kMatrix, (size_t)n * ld * sizeof(T), &dA));
    LAIR_CHECK(pool().get(DevicePool::kRhs, (size_t)n * ldb * sizeof(T), &dB));
    LAIR_CHECK(pool().get(DevicePool::kPivots, (size_t)n * sizeof(int32_t), &dP));
    LAIR_CHECK(upload_matrix<T>(lu, n, n, lu_rs, lu_cs, (T*)dA, ld, DevicePool::kTmpA, s));
    LAIR_CHECK(upload_matrix<T>(b, n, nrhs, b_rs, b_cs, (T*)dB, ldb, DevicePool::kTmpB, s));
    LAIR_CHECK(upload_ipiv32<T>(ipiv, n, (int32_t*)dP, s));
    LAIR_CHECK(getrs_dev<T>(n, nrhs, (const T*)dA, ld, (const int32_t*)dP, (T*)dB, ldb, s));
    LAIR_CHECK(download_matrix<T>(x, n, nrhs, x_rs, x_cs, (const T*)dB, ldb, DevicePool::kTmpB, s));
    return check_fault(s);  // waits for the stream; fails loudly if a device-side wait timed out
}

template <class T>
static int gesv_host(int64_t n, int64_t nrhs, const T* a, int64_t a_rs, int64_t a_cs, const T* b, int64_t b_rs,
                     int64_t b_cs, T* x, int64_t x_rs, int64_t x_cs, int64_t* info) {
    LAIR_REQUIRE(n >= 0 && nrhs >= 0, "gesv: negative dimension");
    LAIR_REQUIRE(info != nullptr, "gesv: info is null");
    *info = -1;
    if (n == 0 || nrhs == 0) return LAIR_B200_OK;
    LAIR_REQUIRE(a && b && x, "gesv: null pointer");
    std::lock_guard<std::mutex> lk(g_call_mu);
    LAIR_CHECK(ensure_init());
    cudaStream_t s = ctx().stream;
    const bool std_layout = is_standard_layout(n, n, a_rs, a_cs);
    const int64_t ld = device_ld(n);
    const int64_t ldb = nrhs;
    void *dA = nullptr, *dB = nullptr, *dP = nullptr, *dI = nullptr;
    LAIR_CHECK(pool().get(DevicePool::kMatrix, (size_t)n * ld * sizeof(T), &dA));
    LAIR_CHECK(pool().get(DevicePool::kRhs, (size_t)n * ldb * sizeof(T), &dB));
    LAIR_CHECK(pool().get(DevicePool::kPivots, (size_t)n * sizeof(int32_t), &dP));
    LAIR_CHECK(pool().get(DevicePool::kInfo, sizeof(int32_t), &dI));
    ColumnFeed feed;
    bool chunked = false;
    LAIR_CHECK(upload_matrix_chunked<T>(a, n, n, a_rs, a_cs, (T*)dA, ld, &feed, &chunked));
    if (!chunked) LAIR_CHECK(upload_matrix<T>(a, n, n, a_rs, a_cs, (T*)dA, ld, DevicePool::kTmpA, s));
    LAIR_CHECK(getrf_dev<T>(n, n, (T*)dA, ld, (int32_t*)dP, (int32_t*)dI, std_layout, s, chunked ? &feed : nullptr));
    // the right-hand sides travel behind the matrix chunks: they are not needed before the solve
    LAIR_CHECK(upload_matrix<T>(b, n, nrhs, b_rs, b_cs, (T*)dB, ldb, DevicePool::kTmpB, s));
    int32_t info32 = -1;
    LAIR_CUDA_CHECK(cudaMemcpyAsync(&info32, dI, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    LAIR_CHECK(check_fault(s));
    *info = info32;
    if (info32 >= 0) return LAIR_B200_OK;  // singular: equation.rs:55-56 returns Err(Value), no solve
    LAIR_CHECK(getrs_dev<T>(n, nrhs, (const T*)dA, ld, (const int32_t*)dP, (T*)dB, ldb, s));
    LAIR_CHECK(download_matrix<T>(x, n, nrhs, x_rs, x_cs, (const T*)dB, ldb, DevicePool::kTmpB, s));
    return check_fault(s);  // waits for the stream; fails loudly if a device-side wait timed out
}

template <class T>
static int getrf_batched_host(int64_t batch, int64_t n, T* a, int32_t* ipiv, int32_t* info) {
    LAIR_REQUIRE(batch >= 0 && n >= 0 && n <= 32, "getrf_batched: need batch >= 0 and 0 <= n <= 32");
    if (batch == 0 || n == 0) return LAIR_B200_OK;
    LAIR_REQUIRE(a && ipiv && info, "getrf_batched: null pointer");
    std::lock_guard<std::mutex> lk(g_call_mu);
    LAIR_CHECK(ensure_init());
    Context& c = ctx();
    cudaStream_t s = c.stream, up = c.copy_stream, down = c.aux_stream;
    void *dA = nullptr, *dP = nullptr, *dI = nullptr;
    const size_t mat_bytes = (size_t)n * n * sizeof(T), piv_bytes = (size_t)n * sizeof(int32_t);
    LAIR_CHECK(pool().get(DevicePool::kMatrix, (size_t)batch * mat_bytes, &dA));
    LAIR_CHECK(pool().get(DevicePool::kPivots, (size_t)batch * piv_bytes, &dP));
    LAIR_CHECK(pool().get(DevicePool::kMisc, (size_t)batch * sizeof(int32_t), &dI));
    // The batch is independent units, so the transfers are pipelined in chunks: chunk i+1 travels host -> device
    // (copy stream) while chunk i is factored (library stream) and chunk i-1 travels back (third stream).  With pinned
    // host memory the call costs max(H2D, D2H) instead of their sum; pageable memory degrades to the driver's staging.
    constexpr int kHalf = Context::kMaxChunks / 2;  // events [0, kHalf): chunk landed; [kHalf, 2 kHalf): chunk factored
    int64_t per = c.opt.batched_chunk;
    if ((batch + per - 1) / per > kHalf) per = (batch + kHalf - 1) / kHalf;
    const int nchunks = (int)((batch + per - 1) / per);
    for (int i = 0; i < 2 * kHalf; ++i)
        if ((i % kHalf) < nchunks && !c.chunk_ev[i]) LAIR_CUDA_CHECK(cudaEventCreateWithFlags(&c.chunk_ev[i], cudaEventDisableTiming));
    // the copy / return streams start after whatever is already queued on the library stream (buffer reuse)
    LAIR_CUDA_CHECK(cudaEventRecord(c.ev[3], s));
    LAIR_CUDA_CHECK(cudaStreamWaitEvent(up, c.ev[3], 0));
    LAIR_CUDA_CHECK(cudaStreamWaitEvent(down, c.ev[3], 0));
    for (int i = 0; i < nchunks; ++i) {
        const int64_t b0 = (int64_t)i * per, nb = (b0 + per <= batch) ? per : (batch - b0);
        char* dAi = (char*)dA + (size_t)b0 * mat_bytes;
        int32_t* dPi = (int32_t*)dP + b0 * n;
        int32_t* dIi = (int32_t*)dI + b0;
        LAIR_CUDA_CHECK(cudaMemcpyAsync(dAi, (const char*)a + (size_t)b0 * mat_bytes, (size_t)nb * mat_bytes, cudaMemcpyHostToDevice, up));
        LAIR_CUDA_CHECK(cudaEventRecord(c.chunk_ev[i], up));
        LAIR_CUDA_CHECK(cudaStreamWaitEvent(s, c.chunk_ev[i], 0));
        LAIR_CHECK(getrf_batched_dev<T>(nb, n, (T*)dAi, dPi, dIi, s));
        LAIR_CUDA_CHECK(cudaEventRecord(c.chunk_ev[kHalf + i], s));
        LAIR_CUDA_CHECK(cudaStreamWaitEvent(down, c.chunk_ev[kHalf + i], 0));
        LAIR_CUDA_CHECK(cudaMemcpyAsync((char*)a + (size_t)b0 * mat_bytes, dAi, (size_t)nb * mat_bytes, cudaMemcpyDeviceToHost, down));
    }
    // Pivots and info (n + 1 int32 per matrix, 1.6 - 3 % of the bytes) return in one piece after the last chunk: callers
    // usually hand in freshly allocated pageable arrays for them, and a pageable D2H blocks the issuing thread until it
    // has finished -- inside the loop that would hold back the next chunk's H2D and serialise the pipeline.
    LAIR_CUDA_CHECK(cudaMemcpyAsync(ipiv, dP, (size_t)batch * piv_bytes, cudaMemcpyDeviceToHost, down));
    LAIR_CUDA_CHECK(cudaMemcpyAsync(info, dI, (size_t)batch * sizeof(int32_t), cudaMemcpyDeviceToHost, down));
    // everything rejoins the library stream before the call returns
    LAIR_CUDA_CHECK(cudaEventRecord(c.ev[3], down));
    LAIR_CUDA_CHECK(cudaStreamWaitEvent(s, c.ev[3], 0));
    LAIR_CUDA_CHECK(cudaStreamSynchronize(s));
    return LAIR_B200_OK;
}

// ---- device-resident factors: lu::Factorized (lu.rs:12-20) behind a handle ---------------------
struct LuHandle {
    uint32_t magic = 0x4c554844u;  // "LUHD"
    int dtype = -1;                // 0 f32, 1 f64, 2 c32, 3 c64
    int64_t m = 0, n = 0, k = 0, ld = 0;
    void* d_lu = nullptr;
    int32_t* d_ipiv = nullptr;
    int32_t info = -1;
};
template <class T> struct DtypeCode;
template <> struct DtypeCode<float> { static constexpr int v = 0; };
template <> struct DtypeCode<double> { static constexpr int v = 1; };
template <> struct DtypeCode<cxf> { static constexpr int v = 2; };
template <> struct DtypeCode<cxd> { static constexpr int v = 3; };

static int lu_check(const LuHandle* h) {
    LAIR_REQUIRE(h != nullptr && h->magic == 0x4c554844u, "lu handle: null or not a handle");
    return LAIR_B200_OK;
}

static void lu_free(LuHandle* h) {
    if (h->d_lu) cudaFree(h->d_lu);
    if (h->d_ipiv) cudaFree(h->d_ipiv);
    h->magic = 0;
    delete h;
}

// Factorized::from (lu.rs:156-171): upload + factor; nothing comes back but `info`.
template <class T>
static int lu_factor_host(int64_t m, int64_t n, const T* a, int64_t rs, int64_t cs, lair_b200_lu_t* out, int64_t* info) {
    LAIR_REQUIRE(m >= 0 && n >= 0, "lu_factor: negative dimension (m=%lld, n=%lld)", (long long)m, (long long)n);
    LAIR_REQUIRE(out != nullptr && info != nullptr, "lu_factor: null output pointer");
    *out = nullptr;
    *info = -1;
    const int64_t k = m < n ? m : n;
    LAIR_REQUIRE(k == 0 || a != nullptr, "lu_factor: null matrix");
    std::lock_guard<std::mutex> lk(g_call_mu);
    LAIR_CHECK(ensure_init());
    LuHandle* h = new (std::nothrow) LuHandle();
    LAIR_REQUIRE(h != nullptr, "lu_factor: out of host memory");
    h->dtype = DtypeCode<T>::v;
    h->m = m;
    h->n = n;
    h->k = k;
    h->ld = device_ld(n);
    if (k > 0) {
        cudaStream_t s = ctx().stream;
        if (cudaMalloc(&h->d_lu, (size_t)m * h->ld * sizeof(T)) != cudaSuccess ||
            cudaMalloc((void**)&h->d_ipiv, (size_t)k * sizeof(int32_t)) != cudaSuccess) {
            cudaGetLastError();
            lu_free(h);
            set_error("lu_factor: device allocation of %zu bytes failed", (size_t)m * (size_t)device_ld(n) * sizeof(T));
            return LAIR_B200_ERR_ALLOC;
        }
        void* dI = nullptr;
        int st = pool().get(DevicePool::kInfo, sizeof(int32_t), &dI);
        const bool std_layout = is_standard_layout(m, n, rs, cs);
        ColumnFeed feed;
        bool chunked = false;
        if (st == LAIR_B200_OK) st = upload_matrix_chunked<T>(a, m, n, rs, cs, (T*)h->d_lu, h->ld, &feed, &chunked);
        if (st == LAIR_B200_OK && !chunked) st = upload_matrix<T>(a, m, n, rs, cs, (T*)h->d_lu, h->ld, DevicePool::kTmpA, s);
        if (st == LAIR_B200_OK) st = getrf_dev<T>(m, n, (T*)h->d_lu, h->ld, h->d_ipiv, (int32_t*)dI, std_layout, s, chunked ? &feed : nullptr);
        int32_t info32 = -1;
        if (st == LAIR_B200_OK && cudaMemcpyAsync(&info32, dI, sizeof(int32_t), cudaMemcpyDeviceToHost, s) != cudaSuccess) {
            set_error("lu_factor: reading info failed: %s", cudaGetErrorString(cudaGetLastError()));
            st = LAIR_B200_ERR_CUDA;
        }
        if (st == LAIR_B200_OK) st = check_fault(s);  // waits for the stream
        if (st != LAIR_B200_OK) {
            cudaStreamSynchronize(s);
            lu_free(h);
            return st;
        }
        h->info = info32;
    }
    *info = h->info;
    *out = reinterpret_cast<lair_b200_lu_t>(h);
    return LAIR_B200_OK;
}

// Factorized::solve (lu.rs:87-98) -> getrs (getrs.rs:12-38), any number of right-hand sides; L\U stays where it is.
template <class T>
static int lu_solve_host(const LuHandle* h, int64_t nrhs, const T* b, int64_t b_rs, int64_t b_cs, T* x, int64_t x_rs, int64_t x_cs) {
    LAIR_REQUIRE(nrhs >= 0, "lu_solve: negative nrhs");
    // getrs.rs:18-20 asks for a.nrows() == p.len() == b.len() and a.ncols() >= p.len(): a WIDE factorization solves with its
    // leading m x m block
    LAIR_REQUIRE(h->m <= h->n, "lu_solve: needs at least as many columns as rows (%lld x %lld); getrs.rs:18-20", (long long)h->m, (long long)h->n);
    const int64_t n = h->m;
    if (n == 0 || nrhs == 0) return LAIR_B200_OK;
    LAIR_REQUIRE(b && x, "lu_solve: null pointer");
    std::lock_guard<std::mutex> lk(g_call_mu);
    LAIR_CHECK(ensure_init());
    cudaStream_t s = ctx().stream;
    const int64_t ldb = nrhs;
    void* dB = nullptr;
    LAIR_CHECK(pool().get(DevicePool::kRhs, (size_t)n * ldb * sizeof(T), &dB));
    LAIR_CHECK(upload_matrix<T>(b, n, nrhs, b_rs, b_cs, (T*)dB, ldb, DevicePool::kTmpB, s));
    LAIR_CHECK(getrs_dev<T>(n, nrhs, (const T*)h->d_lu, h->ld, h->d_ipiv, (T*)dB, ldb, s));
    LAIR_CHECK(download_matrix<T>(x, n, nrhs, x_rs, x_cs, (const T*)dB, ldb, DevicePool::kTmpB, s));
    return check_fault(s);
}

template <class T>
static int lu_factors_host(const LuHandle* h, T* lu, int64_t rs, int64_t cs) {
    if (h->m == 0 || h->n == 0) return LAIR_B200_OK;
    LAIR_REQUIRE(lu != nullptr, "lu_factors: null pointer");
    std::lock_guard<std::mutex> lk(g_call_mu);
    LAIR_CHECK(ensure_init());
    cudaStream_t s = ctx().stream;
    if (h->k == 0) return LAIR_B200_OK;
    LAIR_CHECK(download_matrix<T>(lu, h->m, h->n, rs, cs, (const T*)h->d_lu, h->ld, DevicePool::kTmpA, s));
    LAIR_CUDA_CHECK(cudaStreamSynchronize(s));
    return LAIR_B200_OK;
}

// Factorized::{l, u, p, into_pl} (lu.rs:42-57, 60-72, 28-39, 107-153) from the resident factors.
template <class T>
static int lu_view_host(const LuHandle* h, int view, T* out, int64_t rs, int64_t cs) {
    LAIR_REQUIRE(view >= LAIR_LU_VIEW_L && view <= LAIR_LU_VIEW_PL, "lu_view: unknown view %d", view);
    const int64_t m = h->m, n = h->n, k = h->k;
    const int64_t rows = view == LAIR_LU_VIEW_U ? k : m;
    const int64_t cols = view == LAIR_LU_VIEW_U ? n : (view == LAIR_LU_VIEW_P ? m : k);
    if (rows == 0 || cols == 0) return LAIR_B200_OK;
    LAIR_REQUIRE(out != nullptr, "lu_view: null pointer");
    std::lock_guard<std::mutex> lk(g_call_mu);
    LAIR_CHECK(ensure_init());
    cudaStream_t s = ctx().stream;
    void *dV = nullptr, *dDst = nullptr;
    LAIR_CHECK(pool().get(DevicePool::kRhs, (size_t)rows * cols * sizeof(T), &dV));
    if (view == LAIR_LU_VIEW_P || view == LAIR_LU_VIEW_PL) {
        LAIR_CHECK(pool().get(DevicePool::kMisc, (size_t)m * sizeof(int32_t), &dDst));
        LAIR_CHECK(laswp_follow_dev(m, k, h->d_ipiv, (int32_t*)dDst, s));
    }
    LAIR_CHECK(lu_extract_dev<T>(view, m, n, (const T*)h->d_lu, h->ld, (const int32_t*)dDst, (T*)dV, cols, s));
    LAIR_CHECK(download_matrix<T>(out, rows, cols, rs, cs, (const T*)dV, cols, DevicePool::kTmpB, s));
    LAIR_CUDA_CHECK(cudaStreamSynchronize(s));
    return LAIR_B200_OK;
}

#define INST_DISPATCH(T)                                                                                         \
    template int getrf_dev<T>(int64_t, int64_t, T*, int64_t, int32_t*, int32_t*, bool, cudaStream_t, const ColumnFeed*, RowDrain*); \
    template int getrs_dev<T>(int64_t, int64_t, const T*, int64_t, const int32_t*, T*, int64_t, cudaStream_t);
INST_DISPATCH(float)
INST_DISPATCH(double)
INST_DISPATCH(cxf)
INST_DISPATCH(cxd)

}  // namespace lair

using namespace lair;


// ---- host-pointer QR -------------------------------------------------------
// lapack::geqrf (src/lapack/geqrf.rs:9-30): a (any strides) is overwritten with R and the reflectors, tau gets
// min(m, n) entries.
template <class T>
static int geqrf_host(int64_t m, int64_t n, T* a, int64_t rs, int64_t cs, T* tau) {
    LAIR_REQUIRE(m >= 0 && n >= 0, "geqrf: negative dimension (m=%lld, n=%lld)", (long long)m, (long long)n);
    const int64_t k = m < n ? m : n;
    if (k == 0) return LAIR_B200_OK;
    LAIR_REQUIRE(a != nullptr && tau != nullptr, "geqrf: null pointer");
    std::lock_guard<std::mutex> lk(g_call_mu);
    LAIR_CHECK(ensure_init());
    cudaStream_t s = ctx().stream;
    const int64_t ld = device_ld(n);
    void *dA = nullptr, *dT = nullptr;
    LAIR_CHECK(pool().get(DevicePool::kMatrix, (size_t)m * ld * sizeof(T), &dA));
    LAIR_CHECK(pool().get(DevicePool::kRhs, (size_t)k * sizeof(T), &dT));
    LAIR_CHECK(upload_matrix<T>(a, m, n, rs, cs, (T*)dA, ld, DevicePool::kTmpA, s));
    LAIR_CHECK(geqrf_dev<T>(m, n, (T*)dA, ld, (T*)dT, s));
    LAIR_CHECK(download_matrix<T>(a, m, n, rs, cs, (const T*)dA, ld, DevicePool::kTmpA, s));
    LAIR_CUDA_CHECK(cudaMemcpyAsync(tau, dT, (size_t)k * sizeof(T), cudaMemcpyDeviceToHost, s));
    return check_fault(s);
}

// qr::Factorized::q (src/decomposition/qr.rs:27-59): q (m x m, any strides) from the factored matrix and tau.
template <class T>
static int qr_q_host(int64_t m, int64_t n, const T* qr, int64_t rs, int64_t cs, const T* tau, T* q, int64_t q_rs, int64_t q_cs) {
    LAIR_REQUIRE(m >= 0 && n >= 0, "qr_q: negative dimension");
    if (m == 0) return LAIR_B200_OK;
    const int64_t k = m < n ? m : n;
    LAIR_REQUIRE(q != nullptr && (k == 0 || (qr != nullptr && tau != nullptr)), "qr_q: null pointer");
    std::lock_guard<std::mutex> lk(g_call_mu);
    LAIR_CHECK(ensure_init());
    cudaStream_t s = ctx().stream;
    const int64_t ld = device_ld(n > 0 ? n : 1), ldq = device_ld(m);
    void *dA = nullptr, *dT = nullptr, *dQ = nullptr;
    LAIR_CHECK(pool().get(DevicePool::kMatrix, (size_t)m * ld * sizeof(T), &dA));
    LAIR_CHECK(pool().get(DevicePool::kRhs, (size_t)(k > 0 ? k : 1) * sizeof(T), &dT));
    LAIR_CHECK(pool().get(DevicePool::kTmpB, (size_t)m * ldq * sizeof(T), &dQ));
    if (n > 0) LAIR_CHECK(upload_matrix<T>(qr, m, n, rs, cs, (T*)dA, ld, DevicePool::kTmpA, s));
    if (k > 0) LAIR_CUDA_CHECK(cudaMemcpyAsync(dT, tau, (size_t)k * sizeof(T), cudaMemcpyHostToDevice, s));
    LAIR_CHECK(qr_q_dev<T>(m, n, (const T*)dA, ld, (const T*)dT, (T*)dQ, ldq, s));
    LAIR_CHECK(download_matrix<T>(q, m, m, q_rs, q_cs, (const T*)dQ, ldq, DevicePool::kTmpA, s));
    return check_fault(s);
}

#define DEV_PROLOGUE()        \
    LAIR_CHECK(ensure_init()); \
    cudaStream_t s = (cudaStream_t)stream

extern "C" {

int lair_b200_sgetrf(int64_t m, int64_t n, float* a, int64_t rs, int64_t cs, int64_t* ipiv, int64_t* info) {
    return getrf_host<float>(m, n, a, rs, cs, ipiv, info);
}
int lair_b200_dgetrf(int64_t m, int64_t n, double* a, int64_t rs, int64_t cs, int64_t* ipiv, int64_t* info) {
    return getrf_host<double>(m, n, a, rs, cs, ipiv, info);
}
int lair_b200_cgetrf(int64_t m, int64_t n, void* a, int64_t rs, int64_t cs, int64_t* ipiv, int64_t* info) {
    return getrf_host<cxf>(m, n, (cxf*)a, rs, cs, ipiv, info);
}
int lair_b200_zgetrf(int64_t m, int64_t n, void* a, int64_t rs, int64_t cs, int64_t* ipiv, int64_t* info) {
    return getrf_host<cxd>(m, n, (cxd*)a, rs, cs, ipiv, info);
}

int lair_b200_sgetrs(int64_t n, int64_t nrhs, const float* lu, int64_t lu_rs, int64_t lu_cs, const int64_t* ipiv,
                     const float* b, int64_t b_rs, int64_t b_cs, float* x, int64_t x_rs, int64_t x_cs) {
    return getrs_host<float>(n, nrhs, lu, lu_rs, lu_cs, ipiv, b, b_rs, b_cs, x, x_rs, x_cs);
}
int lair_b200_dgetrs(int64_t n, int64_t nrhs, const double* lu, int64_t lu_rs, int64_t lu_cs, const int64_t* ipiv,
                     const double* b, int64_t b_rs, int64_t b_cs, double* x, int64_t x_rs, int64_t x_cs) {
    return getrs_host<double>(n, nrhs, lu, lu_rs, lu_cs, ipiv, b, b_rs, b_cs, x, x_rs, x_cs);
}
int lair_b200_cgetrs(int64_t n, int64_t nrhs, const void* lu, int64_t lu_rs, int64_t lu_cs, const int64_t* ipiv,
                     const void* b, int64_t b_rs, int64_t b_cs, void* x, int64_t x_rs, int64_t x_cs) {
    return getrs_host<cxf>(n, nrhs, (const cxf*)lu, lu_rs, lu_cs, ipiv, (const cxf*)b, b_rs, b_cs, (cxf*)x, x_rs, x_cs);
}
int lair_b200_zgetrs(int64_t n, int64_t nrhs, const void* lu, int64_t lu_rs, int64_t lu_cs, const int64_t* ipiv,
                     const void* b, int64_t b_rs, int64_t b_cs, void* x, int64_t x_rs, int64_t x_cs) {
    return getrs_host<cxd>(n, nrhs, (const cxd*)lu, lu_rs, lu_cs, ipiv, (const cxd*)b, b_rs, b_cs, (cxd*)x, x_rs, x_cs);
}

int lair_b200_sgesv(int64_t n, int64_t nrhs, const float* a, int64_t a_rs, int64_t a_cs, const float* b, int64_t b_rs,
                    int64_t b_cs, float* x, int64_t x_rs, int64_t x_cs, int64_t* info) {
    return gesv_host<float>(n, nrhs, a, a_rs, a_cs, b, b_rs, b_cs, x, x_rs, x_cs, info);
}
int lair_b200_dgesv(int64_t n, int64_t nrhs, const double* a, int64_t a_rs, int64_t a_cs, const double* b, int64_t b_rs,
                    int64_t b_cs, double* x, int64_t x_rs, int64_t x_cs, int64_t* info) {
    return gesv_host<double>(n, nrhs, a, a_rs, a_cs, b, b_rs, b_cs, x, x_rs, x_cs, info);
}

int lair_b200_sgetrf_batched(int64_t batch, int64_t n, float* a, int32_t* ipiv, int32_t* info) {
    return getrf_batched_host<float>(batch, n, a, ipiv, info);
}
int lair_b200_dgetrf_batched(int64_t batch, int64_t n, double* a, int32_t* ipiv, int32_t* info) {
    return getrf_batched_host<double>(batch, n, a, ipiv, info);
}


// ---- device-resident factors -------------------------------------------------------------------
int lair_b200_slu_factor(int64_t m, int64_t n, const float* a, int64_t rs, int64_t cs, lair_b200_lu_t* handle, int64_t* info) {
    return lu_factor_host<float>(m, n, a, rs, cs, handle, info);
}
int lair_b200_dlu_factor(int64_t m, int64_t n, const double* a, int64_t rs, int64_t cs, lair_b200_lu_t* handle, int64_t* info) {
    return lu_factor_host<double>(m, n, a, rs, cs, handle, info);
}
int lair_b200_clu_factor(int64_t m, int64_t n, const void* a, int64_t rs, int64_t cs, lair_b200_lu_t* handle, int64_t* info) {
    return lu_factor_host<cxf>(m, n, (const cxf*)a, rs, cs, handle, info);
}
int lair_b200_zlu_factor(int64_t m, int64_t n, const void* a, int64_t rs, int64_t cs, lair_b200_lu_t* handle, int64_t* info) {
    return lu_factor_host<cxd>(m, n, (const cxd*)a, rs, cs, handle, info);
}
int lair_b200_lu_solve(lair_b200_lu_t handle, int64_t nrhs, const void* b, int64_t b_rs, int64_t b_cs, void* x, int64_t x_rs, int64_t x_cs) {
    const LuHandle* h = reinterpret_cast<const LuHandle*>(handle);
    LAIR_CHECK(lu_check(h));
    switch (h->dtype) {
        case 0: return lu_solve_host<float>(h, nrhs, (const float*)b, b_rs, b_cs, (float*)x, x_rs, x_cs);
        case 1: return lu_solve_host<double>(h, nrhs, (const double*)b, b_rs, b_cs, (double*)x, x_rs, x_cs);
        case 2: return lu_solve_host<cxf>(h, nrhs, (const cxf*)b, b_rs, b_cs, (cxf*)x, x_rs, x_cs);
        default: return lu_solve_host<cxd>(h, nrhs, (const cxd*)b, b_rs, b_cs, (cxd*)x, x_rs, x_cs);
    }
}
int lair_b200_lu_pivots(lair_b200_lu_t handle, int64_t* ipiv) {
    const LuHandle* h = reinterpret_cast<const LuHandle*>(handle);
    LAIR_CHECK(lu_check(h));
    if (h->k == 0) return LAIR_B200_OK;
    LAIR_REQUIRE(ipiv != nullptr, "lu_pivots: null pointer");
    std::lock_guard<std::mutex> lk(g_call_mu);
    LAIR_CHECK(ensure_init());
    LAIR_CHECK(download_ipiv64(ipiv, h->d_ipiv, h->k, DevicePool::kPivots64, ctx().stream));
    LAIR_CUDA_CHECK(cudaStreamSynchronize(ctx().stream));
    return LAIR_B200_OK;
}
int lair_b200_lu_factors(lair_b200_lu_t handle, void* lu, int64_t rs, int64_t cs) {
    const LuHandle* h = reinterpret_cast<const LuHandle*>(handle);
    LAIR_CHECK(lu_check(h));
    switch (h->dtype) {
        case 0: return lu_factors_host<float>(h, (float*)lu, rs, cs);
        case 1: return lu_factors_host<double>(h, (double*)lu, rs, cs);
        case 2: return lu_factors_host<cxf>(h, (cxf*)lu, rs, cs);
        default: return lu_factors_host<cxd>(h, (cxd*)lu, rs, cs);
    }
}
int lair_b200_lu_view(lair_b200_lu_t handle, int view, void* out, int64_t rs, int64_t cs) {
    const LuHandle* h = reinterpret_cast<const LuHandle*>(handle);
    LAIR_CHECK(lu_check(h));
    switch (h->dtype) {
        case 0: return lu_view_host<float>(h, view, (float*)out, rs, cs);
        case 1: return lu_view_host<double>(h, view, (double*)out, rs, cs);
        case 2: return lu_view_host<cxf>(h, view, (cxf*)out, rs, cs);
        default: return lu_view_host<cxd>(h, view, (cxd*)out, rs, cs);
    }
}
int lair_b200_lu_shape(lair_b200_lu_t handle, int64_t* m, int64_t* n, int* dtype, int64_t* info) {
    const LuHandle* h = reinterpret_cast<const LuHandle*>(handle);
    LAIR_CHECK(lu_check(h));
    if (m) *m = h->m;
    if (n) *n = h->n;
    if (dtype) *dtype = h->dtype;
    if (info) *info = h->info;
    return LAIR_B200_OK;
}
int lair_b200_lu_destroy(lair_b200_lu_t handle) {
    LuHandle* h = reinterpret_cast<LuHandle*>(handle);
    if (h == nullptr) return LAIR_B200_OK;
    LAIR_CHECK(lu_check(h));
    std::lock_guard<std::mutex> lk(g_call_mu);
    if (ctx().stream) cudaStreamSynchronize(ctx().stream);
    lu_free(h);
    return LAIR_B200_OK;
}

// ---- device-resident -------------------------------------------------------------------------
int lair_b200_sgetrf_dev(int64_t m, int64_t n, float* d_a, int64_t lda, int32_t* d_ipiv, int32_t* d_info, void* stream) {
    DEV_PROLOGUE();
    LAIR_REQUIRE(m >= 0 && n >= 0 && lda >= n, "getrf_dev: bad shape");
    return getrf_dev<float>(m, n, d_a, lda, d_ipiv, d_info, true, s);
}
int lair_b200_dgetrf_dev(int64_t m, int64_t n, double* d_a, int64_t lda, int32_t* d_ipiv, int32_t* d_info, void* stream) {
    DEV_PROLOGUE();
    LAIR_REQUIRE(m >= 0 && n >= 0 && lda >= n, "getrf_dev: bad shape");
    return getrf_dev<double>(m, n, d_a, lda, d_ipiv, d_info, true, s);
}
// complex: interleaved (re, im) elements, Complex<f32> / Complex<f64> as num-complex lays them out
int lair_b200_cgetrf_dev(int64_t m, int64_t n, void* d_a, int64_t lda, int32_t* d_ipiv, int32_t* d_info, void* stream) {
    DEV_PROLOGUE();
    LAIR_REQUIRE(m >= 0 && n >= 0 && lda >= n, "getrf_dev: bad shape");
    return getrf_dev<cxf>(m, n, static_cast<cxf*>(d_a), lda, d_ipiv, d_info, true, s);
}
int lair_b200_zgetrf_dev(int64_t m, int64_t n, void* d_a, int64_t lda, int32_t* d_ipiv, int32_t* d_info, void* stream) {
    DEV_PROLOGUE();
    LAIR_REQUIRE(m >= 0 && n >= 0 && lda >= n, "getrf_dev: bad shape");
    return getrf_dev<cxd>(m, n, static_cast<cxd*>(d_a), lda, d_ipiv, d_info, true, s);
}
int lair_b200_sgetrs_dev(int64_t n, int64_t nrhs, const float* d_lu, int64_t lda, const int32_t* d_ipiv, float* d_b,
                         int64_t ldb, void* stream) {
    DEV_PROLOGUE();
    return getrs_dev<float>(n, nrhs, d_lu, lda, d_ipiv, d_b, ldb, s);
}
int lair_b200_dgetrs_dev(int64_t n, int64_t nrhs, const double* d_lu, int64_t lda, const int32_t* d_ipiv, double* d_b,
                         int64_t ldb, void* stream) {
    DEV_PROLOGUE();
    return getrs_dev<double>(n, nrhs, d_lu, lda, d_ipiv, d_b, ldb, s);
}
int lair_b200_sgetrf_batched_dev(int64_t batch, int64_t n, float* d_a, int32_t* d_ipiv, int32_t* d_info, void* stream) {
    DEV_PROLOGUE();
    return getrf_batched_dev<float>(batch, n, d_a, d_ipiv, d_info, s);
}
int lair_b200_dgetrf_batched_dev(int64_t batch, int64_t n, double* d_a, int32_t* d_ipiv, int32_t* d_info, void* stream) {
    DEV_PROLOGUE();
    return getrf_batched_dev<double>(batch, n, d_a, d_ipiv, d_info, s);
}

int lair_b200_dlaswp_dev(int64_t ncols, double* d_a, int64_t lda, int64_t k0, int64_t k1, const int32_t* d_ipiv, void* stream) {
    DEV_PROLOGUE();
    return laswp_dev<double>(ncols, d_a, lda, k0, k1, d_ipiv, s);
}
int lair_b200_slaswp_dev(int64_t ncols, float* d_a, int64_t lda, int64_t k0, int64_t k1, const int32_t* d_ipiv, void* stream) {
    DEV_PROLOGUE();
    return laswp_dev<float>(ncols, d_a, lda, k0, k1, d_ipiv, s);
}
int lair_b200_dtrsm_dev(int64_t k, int64_t ncols, const double* d_l, int64_t ldl, double* d_b, int64_t ldb, void* stream) {
    DEV_PROLOGUE();
    return trsm_lower_unit_dev<double>(k, ncols, d_l, ldl, d_b, ldb, s);
}
int lair_b200_strsm_dev(int64_t k, int64_t ncols, const float* d_l, int64_t ldl, float* d_b, int64_t ldb, void* stream) {
    DEV_PROLOGUE();
    return trsm_lower_unit_dev<float>(k, ncols, d_l, ldl, d_b, ldb, s);
}
int lair_b200_dgemm_minus_dev(int64_t m, int64_t n, int64_t k, const double* d_a, int64_t lda, const double* d_b,
                              int64_t ldb, double* d_c, int64_t ldc, void* stream) {
    DEV_PROLOGUE();
    return gemm_minus_dev<double>(m, n, k, d_a, lda, d_b, ldb, d_c, ldc, s);
}
int lair_b200_sgemm_minus_dev(int64_t m, int64_t n, int64_t k, const float* d_a, int64_t lda, const float* d_b,
                              int64_t ldb, float* d_c, int64_t ldc, void* stream) {
    DEV_PROLOGUE();
    return gemm_minus_dev<float>(m, n, k, d_a, lda, d_b, ldb, d_c, ldc, s);
}

// ---- QR ------------------------------------------------------------------------------------------
#define LAIR_QR_ENTRY(P, T, CT)                                                                                          \
    int lair_b200_##P##geqrf(int64_t m, int64_t n, CT* a, int64_t rs, int64_t cs, CT* tau) {                              \
        return geqrf_host<T>(m, n, (T*)a, rs, cs, (T*)tau);                                                              \
    }                                                                                                                    \
    int lair_b200_##P##qr_q(int64_t m, int64_t n, const CT* qr, int64_t rs, int64_t cs, const CT* tau, CT* q, int64_t q_rs, \
                            int64_t q_cs) {                                                                              \
        return qr_q_host<T>(m, n, (const T*)qr, rs, cs, (const T*)tau, (T*)q, q_rs, q_cs);                               \
    }                                                                                                                    \
    int lair_b200_##P##geqrf_dev(int64_t m, int64_t n, CT* d_a, int64_t lda, CT* d_tau, void* stream) {                   \
        DEV_PROLOGUE();                                                                                                  \
        return geqrf_dev<T>(m, n, (T*)d_a, lda, (T*)d_tau, s);                                                           \
    }
LAIR_QR_ENTRY(s, float, float)
LAIR_QR_ENTRY(d, double, double)
LAIR_QR_ENTRY(c, cxf, void)
LAIR_QR_ENTRY(z, cxd, void)
#undef LAIR_QR_ENTRY

}  // extern "C"
