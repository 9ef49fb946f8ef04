// TEST INFRASTRUCTURE ONLY -- extern "C" face of the CPU oracle (see
// lair_oracle.hpp for the parity-pin statement and the reference citations).
// Loaded through ctypes by tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs; never by the product path.
#include "lair_oracle.hpp"

#include <chrono>
#include <cstring>

using namespace lair_oracle;

#define ORACLE_FOR_TYPE(P, T)                                                                      \
    /* lapack::getrf: returns last zero-pivot step or -1 (src/lapack/getrf.rs:12-27) */           \
    extern "C" int64_t oracle_##P##getrf(int64_t m, int64_t n, void* a, int64_t rs, int64_t cs,    \
                                         uint64_t* piv) {                                          \
        return getrf<T>((T*)a, (size_t)m, (size_t)n, (ptrdiff_t)rs, (ptrdiff_t)cs, (size_t*)piv);  \
    }                                                                                              \
    /* force one variant regardless of layout: which=0 row-major, 1 col-major */                   \
    extern "C" int64_t oracle_##P##getrf_variant(int which, int64_t m, int64_t n, void* a,         \
                                                 int64_t rs, int64_t cs, uint64_t* piv) {          \
        size_t dm = (size_t)std::min(m, n);                                                        \
        for (size_t i = 0; i < dm; ++i) piv[i] = i;                                                \
        if (which == 0)                                                                            \
            return getrf_row_major<T>((T*)a, (size_t)m, (size_t)n, rs, cs, (size_t*)piv);          \
        return getrf_col_major<T>((T*)a, (size_t)m, (size_t)n, rs, cs, (size_t*)piv);              \
    }                                                                                              \
    /* getrf_recursive: piv has m entries; returns -1 Ok, else Singular(row) */                    \
    extern "C" int64_t oracle_##P##getrf_recursive(int64_t m, int64_t n, void* a, int64_t rs,      \
                                                   int64_t cs, uint64_t* piv) {                    \
        for (int64_t i = 0; i < m; ++i) piv[i] = 0;                                                \
        size_t row = 0;                                                                            \
        bool ok = recursive_inner<T>((T*)a, (size_t)m, (size_t)n, rs, cs, (size_t*)piv, &row);     \
        return ok ? -1 : (int64_t)row;                                                             \
    }                                                                                              \
    extern "C" void oracle_##P##getrs(int64_t n, const void* a, int64_t rs, int64_t cs,            \
                                      const uint64_t* piv, const void* b, int64_t incb, void* x) { \
        getrs<T>((const T*)a, (size_t)n, rs, cs, (const size_t*)piv, (const T*)b, incb, (T*)x);    \
    }                                                                                              \
    extern "C" void oracle_##P##laswp(int64_t ncols, void* a, int64_t rs, int64_t cs,              \
                                      int64_t begin, const uint64_t* piv, int64_t npiv) {          \
        laswp<T>((size_t)ncols, (T*)a, rs, cs, (size_t)begin, (const size_t*)piv, (size_t)npiv);   \
    }                                                                                              \
    extern "C" void oracle_##P##trsm(const void* a, int64_t a_rs, int64_t a_cs, void* b,           \
                                     int64_t b_rows, int64_t b_cols, int64_t b_rs, int64_t b_cs) { \
        trsm<T>((const T*)a, a_rs, a_cs, (T*)b, (size_t)b_rows, (size_t)b_cols, b_rs, b_cs);       \
    }                                                                                              \
    extern "C" void oracle_##P##gemm_minus(const void* a, int64_t m, int64_t k, int64_t a_rs,      \
                                           int64_t a_cs, const void* b, int64_t n, int64_t b_rs,   \
                                           int64_t b_cs, void* c, int64_t c_rs, int64_t c_cs) {    \
        gemm<T>(-ScalarTraits<T>::one(), (const T*)a, (size_t)m, (size_t)k, a_rs, a_cs,            \
                (const T*)b, (size_t)n, b_rs, b_cs, (T*)c, c_rs, c_cs);                            \
    }                                                                                              \
    extern "C" void oracle_##P##into_pl(int64_t m, int64_t n, void* lu, int64_t rs, int64_t cs,    \
                                        const uint64_t* piv, int64_t npiv) {                       \
        std::vector<size_t> p(piv, piv + npiv);                                                    \
        into_pl<T>((T*)lu, (size_t)m, (size_t)n, rs, cs, p);                                       \
    }                                                                                              \
    /* batched helper for timing: `batch` contiguous row-major n x n matrices */                   \
    extern "C" void oracle_##P##getrf_batched(int64_t batch, int64_t n, void* a, uint64_t* piv,    \
                                              int64_t* info) {                                     \
        for (int64_t b = 0; b < batch; ++b) {                                                      \
            info[b] = getrf<T>((T*)a + b * n * n, (size_t)n, (size_t)n, (ptrdiff_t)n, 1,           \
                               (size_t*)piv + b * n);                                              \
        }                                                                                          \
    }

#define ORACLE_QR_FOR_TYPE(P, T)                                                                   \
    /* lapack::geqrf (src/lapack/geqrf.rs:9-30) */                                                 \
    extern "C" void oracle_##P##geqrf(int64_t m, int64_t n, void* a, int64_t rs, int64_t cs, void* tau) { \
        geqrf<T>((T*)a, (size_t)m, (size_t)n, rs, cs, (T*)tau);                                    \
    }                                                                                              \
    /* qr::Factorized::q (src/decomposition/qr.rs:27-59): q = m x m row-major */                   \
    extern "C" void oracle_##P##qr_q(int64_t m, int64_t n, const void* qr, int64_t rs, int64_t cs, \
                                     const void* tau, void* q) {                                   \
        qr_q<T>((const T*)qr, (size_t)m, (size_t)n, rs, cs, (const T*)tau, (T*)q);                 \
    }                                                                                              \
    /* lapack::larfg (src/lapack/larfg.rs:9-42): alpha_beta in: alpha, out: (beta, 0) */           \
    extern "C" void oracle_##P##larfg(void* alpha_beta, int64_t n, void* x, int64_t inc, void* tau) { \
        RealOf<T> beta;                                                                            \
        T t;                                                                                       \
        larfg<T>(*(T*)alpha_beta, (size_t)n, (T*)x, inc, &beta, &t);                               \
        *(T*)alpha_beta = from_real(beta, (T*)nullptr);                                            \
        *(T*)tau = t;                                                                              \
    }
ORACLE_QR_FOR_TYPE(s, float)
ORACLE_QR_FOR_TYPE(d, double)
ORACLE_QR_FOR_TYPE(c, Cx<float>)
ORACLE_QR_FOR_TYPE(z, Cx<double>)

ORACLE_FOR_TYPE(s, float)
ORACLE_FOR_TYPE(d, double)
ORACLE_FOR_TYPE(c, Cx<float>)
ORACLE_FOR_TYPE(z, Cx<double>)

// blas::iamax (src/blas/iamax.rs:6-21)
extern "C" int64_t oracle_diamax(int64_t n, const double* x, int64_t incx, double* max_val) {
    size_t idx;
    iamax<double>((size_t)n, x, incx, &idx, max_val);
    return (int64_t)idx;
}
extern "C" int64_t oracle_siamax(int64_t n, const float* x, int64_t incx, float* max_val) {
    size_t idx;
    iamax<float>((size_t)n, x, incx, &idx, max_val);
    return (int64_t)idx;
}
extern "C" int64_t oracle_ziamax(int64_t n, const void* x, int64_t incx, double* max_val) {
    size_t idx;
    iamax<Cx<double>>((size_t)n, (const Cx<double>*)x, incx, &idx, max_val);
    return (int64_t)idx;
}

// Timed leg for bench.py: factor (and optionally solve nrhs right-hand sides one
// by one, as the reference's single-RHS getrs forces) a row-major f64 system on
// ONE core; returns seconds.  `a` is overwritten with L\U.
extern "C" double oracle_time_dgetrf_dgetrs(int64_t n, double* a, uint64_t* piv, int64_t nrhs,
                                            const double* b /* n x nrhs row-major */,
                                            double* x /* n x nrhs row-major */,
                                            double* secs_getrf, double* secs_getrs) {
    auto t0 = std::chrono::steady_clock::now();
    getrf<double>(a, (size_t)n, (size_t)n, (ptrdiff_t)n, 1, (size_t*)piv);
    auto t1 = std::chrono::steady_clock::now();
    std::vector<double> xc((size_t)n);
    for (int64_t r = 0; r < nrhs; ++r) {
        getrs<double>(a, (size_t)n, (ptrdiff_t)n, 1, (const size_t*)piv, b + r, (ptrdiff_t)nrhs,
                      xc.data());
        for (int64_t i = 0; i < n; ++i) x[i * nrhs + r] = xc[(size_t)i];
    }
    auto t2 = std::chrono::steady_clock::now();
    double g = std::chrono::duration<double>(t1 - t0).count();
    double s = std::chrono::duration<double>(t2 - t1).count();
    if (secs_getrf) *secs_getrf = g;
    if (secs_getrs) *secs_getrs = s;
    return g + s;
}
