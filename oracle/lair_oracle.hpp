// TEST INFRASTRUCTURE ONLY -- CPU oracle for the lair LU path.
//
// A literal C++ restatement of the reference algorithm (vinesystems/lair v0.8.0,
// pure Rust).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this; the product (lair_b200/) never does and
// fails loudly when its CUDA library is missing.
//
// Parity pin: no Rust toolchain exists in this image, so the reference itself
// cannot be run.  The oracle is pinned against every golden vector the
// reference's own unit tests hold for this path (tests/golden/lair_golden.json,
// transcribed from src/lapack/getrf.rs:343-524, src/lapack/getrs.rs:47-78,
// src/blas/iamax.rs:27-34, src/decomposition/lu.rs:182-312,
// src/equation.rs:21-31) by tests/test_oracle_golden.py.
//
// Build: g++ -O2 -ffp-contract=off (Rust never contracts a*b-c into an FMA, so
// every `a -= l * u` below rounds twice exactly like the reference).
//
// Third-party arithmetic restated here: num-complex 0.4 (Cargo.toml:24, not
// vendored in /root/reference): Complex mul / div use the plain textbook
// formulas (`re = a.re*b.re - a.im*b.im`, division through `norm_sqr`), not
// C99 Annex-G recovery -- so we carry our own Cx<T> instead of std::complex.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <algorithm>
#include <vector>
#include <limits>

namespace lair_oracle {

// ---- scalar layer (src/scalar.rs:332-407) -----------------------------------
template <class T>
struct Cx {
    T re, im;
};
template <class T> inline Cx<T> operator+(Cx<T> a, Cx<T> b) { return {a.re + b.re, a.im + b.im}; }
template <class T> inline Cx<T> operator-(Cx<T> a, Cx<T> b) { return {a.re - b.re, a.im - b.im}; }
template <class T> inline Cx<T> operator-(Cx<T> a) { return {-a.re, -a.im}; }
// num-complex 0.4 `impl Mul for Complex`.
template <class T> inline Cx<T> operator*(Cx<T> a, Cx<T> b) {
    T re = a.re * b.re - a.im * b.im;
    T im = a.re * b.im + a.im * b.re;
    return {re, im};
}
// num-complex 0.4 `impl Div for Complex`.
template <class T> inline Cx<T> operator/(Cx<T> a, Cx<T> b) {
    T norm_sqr = b.re * b.re + b.im * b.im;
    T re = a.re * b.re + a.im * b.im;
    T im = a.im * b.re - a.re * b.im;
    return {re / norm_sqr, im / norm_sqr};
}
template <class T> inline bool operator==(Cx<T> a, Cx<T> b) { return a.re == b.re && a.im == b.im; }

template <class A> struct ScalarTraits {
    using Real = A;
    static A zero() { return A(0); }
    static A one() { return A(1); }
    static Real re(A x) { return x; }
    static Real im(A) { return Real(0); }
};
template <class T> struct ScalarTraits<Cx<T>> {
    using Real = T;
    static Cx<T> zero() { return {T(0), T(0)}; }
    static Cx<T> one() { return {T(1), T(0)}; }
    static Real re(Cx<T> x) { return x.re; }
    static Real im(Cx<T> x) { return x.im; }
};
template <class A> using RealOf = typename ScalarTraits<A>::Real;

// Real::sfmin (src/scalar.rs:400-402) = min_positive_value.
template <class R> inline R sfmin() { return std::numeric_limits<R>::min(); }

// ---- blas::iamax (src/blas/iamax.rs:6-21) -----------------------------------
// First index of max |re|+|im|; strict `>` so the first maximum wins and a NaN
// never wins; all-zero / all-NaN input returns (0, 0).
template <class A>
inline void iamax(size_t n, const A* x, ptrdiff_t incx, size_t* max_idx_out, RealOf<A>* max_val_out) {
    using R = RealOf<A>;
    R max_val = R(0);
    size_t max_idx = 0;
    for (size_t i = 0; i < n; ++i) {
        const A* elem = x + incx * (ptrdiff_t)i;
        R val = std::fabs(ScalarTraits<A>::re(*elem)) + std::fabs(ScalarTraits<A>::im(*elem));
        if (val > max_val) {
            max_val = val;
            max_idx = i;
        }
    }
    *max_idx_out = max_idx;
    *max_val_out = max_val;
}

// ---- swap_rows (src/lapack/getrf.rs:324-333) --------------------------------
template <class A>
inline void swap_rows(size_t n, A* row1, A* row2, ptrdiff_t stride) {
    while (n > 0) {
        A tmp = *row1;
        *row1 = *row2;
        *row2 = tmp;
        row1 += stride;
        row2 += stride;
        --n;
    }
}

// ---- lapack::laswp (src/lapack/laswp.rs:11-40) ------------------------------
// Sequential interchanges piv[begin..]; skips i == p; `ncols` must be >= 1
// (the reference's loop body runs once before testing n == 1).
template <class T>
inline void laswp(size_t ncols, T* a, ptrdiff_t row_stride, ptrdiff_t col_stride, size_t begin,
                  const size_t* piv, size_t npiv) {
    for (size_t i = begin; i < npiv; ++i) {
        size_t p = piv[i];
        if (i == p) continue;
        size_t n = ncols;
        T* row1 = a + (ptrdiff_t)i * row_stride;
        T* row2 = a + (ptrdiff_t)p * row_stride;
        for (;;) {
            T tmp = *row1;
            *row1 = *row2;
            *row2 = tmp;
            if (n == 1) break;
            row1 += col_stride;
            row2 += col_stride;
            --n;
        }
    }
}

// ---- blas::dot (src/blas/dot.rs:39-133) -------------------------------------
// Four interleaved accumulators, combined ((s0+s1)+s2)+s3, then up to three
// tail terms; the contiguous and strided bodies share this order.
template <class T>
inline T dot(size_t n, const T* x, ptrdiff_t inc_x, const T* y, ptrdiff_t inc_y) {
    size_t remaining = n;
    T sum;
    if (remaining >= 4) {
        T s0 = ScalarTraits<T>::zero(), s1 = s0, s2 = s0, s3 = s0;
        for (;;) {
            s0 = s0 + x[0] * y[0];
            s1 = s1 + x[inc_x] * y[inc_y];
            s2 = s2 + x[inc_x * 2] * y[inc_y * 2];
            s3 = s3 + x[inc_x * 3] * y[inc_y * 3];
            x += inc_x * 4;
            y += inc_y * 4;
            remaining -= 4;
            if (remaining < 4) break;
        }
        sum = s0 + s1 + s2 + s3;
    } else {
        sum = ScalarTraits<T>::zero();
    }
    if (remaining > 0) {
        sum = sum + x[0] * y[0];
        x += inc_x;
        y += inc_y;
    }
    if (remaining > 1) {
        sum = sum + x[0] * y[0];
        x += inc_x;
        y += inc_y;
    }
    if (remaining > 2) {
        sum = sum + x[0] * y[0];
    }
    return sum;
}

// ---- blas::gemv::notrans (src/blas/gemv.rs:6-48) ----------------------------
// y = alpha*a*x + beta*y as column axpys: for each column, alpha_x = alpha*x[k],
// y[r] += alpha_x * a[r,k].
template <class A>
inline void gemv_notrans(A alpha, const A* a, size_t nrows, size_t ncols, ptrdiff_t a_rs,
                         ptrdiff_t a_cs, const A* x, ptrdiff_t incx, A beta, A* y, ptrdiff_t incy) {
    const A zero = ScalarTraits<A>::zero(), one = ScalarTraits<A>::one();
    if (beta == zero) {
        for (size_t r = 0; r < nrows; ++r) y[(ptrdiff_t)r * incy] = zero;
    } else if (!(beta == one)) {
        for (size_t r = 0; r < nrows; ++r) y[(ptrdiff_t)r * incy] = y[(ptrdiff_t)r * incy] * beta;
    }
    if (alpha == zero) return;
    for (size_t k = 0; k < ncols; ++k) {
        A alpha_x = alpha * x[(ptrdiff_t)k * incx];
        const A* a_elem = a + (ptrdiff_t)k * a_cs;
        A* y_elem = y;
        for (size_t r = 0; r < nrows; ++r) {
            *y_elem = *y_elem + alpha_x * *a_elem;
            a_elem += a_rs;
            y_elem += incy;
        }
    }
}

// ---- blas::trsm (src/blas/trsm.rs:6-22) -------------------------------------
// B <- L^-1 B, L unit lower from (a, strides); skips exactly-zero b[k,j].
template <class A>
inline void trsm(const A* a, ptrdiff_t a_rs, ptrdiff_t a_cs, A* b, size_t b_rows, size_t b_cols,
                 ptrdiff_t b_rs, ptrdiff_t b_cs) {
    for (size_t j = 0; j < b_cols; ++j) {
        for (size_t k = 0; k < b_rows; ++k) {
            A bkj = b[(ptrdiff_t)k * b_rs + (ptrdiff_t)j * b_cs];
            if (bkj == ScalarTraits<A>::zero()) continue;
            for (size_t i = k + 1; i < b_rows; ++i) {
                A prod = b[(ptrdiff_t)k * b_rs + (ptrdiff_t)j * b_cs] *
                         a[a_rs * (ptrdiff_t)i + a_cs * (ptrdiff_t)k];
                A* bij = &b[(ptrdiff_t)i * b_rs + (ptrdiff_t)j * b_cs];
                *bij = *bij - prod;
            }
        }
    }
}

// ---- blas::gemm (src/blas/gemm.rs:6-32), conj flags false on the LU path ----
// c[i,j] += alpha * fold(0, sum + a[i,k]*b[k,j]).
template <class A>
inline void gemm(A alpha, const A* a, size_t m, size_t k, ptrdiff_t a_rs, ptrdiff_t a_cs,
                 const A* b, size_t n, ptrdiff_t b_rs, ptrdiff_t b_cs, A* c, ptrdiff_t c_rs,
                 ptrdiff_t c_cs) {
    for (size_t i = 0; i < m; ++i) {
        for (size_t j = 0; j < n; ++j) {
            A sum = ScalarTraits<A>::zero();
            for (size_t kk = 0; kk < k; ++kk) {
                sum = sum + a[(ptrdiff_t)i * a_rs + (ptrdiff_t)kk * a_cs] *
                                b[(ptrdiff_t)kk * b_rs + (ptrdiff_t)j * b_cs];
            }
            A* cij = &c[(ptrdiff_t)i * c_rs + (ptrdiff_t)j * c_cs];
            *cij = *cij + alpha * sum;
        }
    }
}

// ---- getrf_row_major (src/lapack/getrf.rs:46-120) ---------------------------
// Right-looking unblocked LU; returns the LAST zero-pivot step or -1.
template <class A>
inline int64_t getrf_row_major(A* a_ptr, size_t nrows, size_t ncols, ptrdiff_t row_stride,
                               ptrdiff_t col_stride, size_t* pivots) {
    using R = RealOf<A>;
    int64_t singular_row = -1;
    size_t dim_min = std::min(nrows, ncols);
    for (size_t ul = 0; ul < dim_min; ++ul) {
        size_t max_idx;
        R max_val;
        iamax<A>(nrows - ul, a_ptr + row_stride * (ptrdiff_t)ul + col_stride * (ptrdiff_t)ul,
                 row_stride, &max_idx, &max_val);
        if (max_idx != 0) {
            size_t max_row = max_idx + ul;
            pivots[ul] = max_row;
            swap_rows(ncols, a_ptr + (ptrdiff_t)ul * row_stride,
                      a_ptr + (ptrdiff_t)max_row * row_stride, col_stride);
        }
        if (max_val == R(0)) {
            singular_row = (int64_t)(max_idx + ul);
        } else {
            A* diag = a_ptr + row_stride * (ptrdiff_t)ul + col_stride * (ptrdiff_t)ul;
            A pivot_recip = ScalarTraits<A>::one() / *diag;
            A* row_j = diag;
            if (col_stride == 1) {
                for (ptrdiff_t j = 1; j < (ptrdiff_t)(nrows - ul); ++j) {
                    row_j += row_stride;
                    *row_j = *row_j * pivot_recip;
                    A ratio = *row_j;
                    A* row_i = a_ptr + row_stride * (ptrdiff_t)ul + (ptrdiff_t)ul;
                    for (size_t c = ul + 1; c < ncols; ++c) {
                        row_i += 1;
                        A elem = ratio * *row_i;
                        A* t = row_i + j * row_stride;
                        *t = *t - elem;
                    }
                }
            } else if (row_stride == 1) {
                for (ptrdiff_t j = 1; j < (ptrdiff_t)(nrows - ul); ++j) {
                    row_j += 1;
                    *row_j = *row_j * pivot_recip;
                    A ratio = *row_j;
                    A* row_i = a_ptr + (ptrdiff_t)ul + col_stride * (ptrdiff_t)ul;
                    for (size_t c = ul + 1; c < ncols; ++c) {
                        row_i += col_stride;
                        A elem = ratio * *row_i;
                        A* t = row_i + j;
                        *t = *t - elem;
                    }
                }
            } else {
                for (ptrdiff_t j = 1; j < (ptrdiff_t)(nrows - ul); ++j) {
                    row_j += row_stride;
                    *row_j = *row_j * pivot_recip;
                    A ratio = *row_j;
                    A* row_i = diag;
                    for (size_t c = ul + 1; c < ncols; ++c) {
                        row_i += col_stride;
                        A elem = ratio * *row_i;
                        A* t = row_i + j * row_stride;
                        *t = *t - elem;
                    }
                }
            }
        }
    }
    return singular_row;
}

// ---- getrf_col_major (src/lapack/getrf.rs:128-213) --------------------------
// Left-looking column sweep used for EVERY non-standard layout.
template <class A>
inline int64_t getrf_col_major(A* a_ptr, size_t nrows, size_t ncols, ptrdiff_t row_stride,
                               ptrdiff_t col_stride, size_t* pivots) {
    using R = RealOf<A>;
    int64_t singular_row = -1;
    for (size_t ul = 0; ul < ncols; ++ul) {
        A* col = a_ptr + (ptrdiff_t)ul * col_stride;  // column `ul`, element stride row_stride
        size_t n_upper_rows = std::min(ul, nrows);
        for (size_t i = 0; i < n_upper_rows; ++i) {
            size_t ip = pivots[i];
            if (ip != i) std::swap(col[(ptrdiff_t)i * row_stride], col[(ptrdiff_t)ip * row_stride]);
        }
        for (size_t i = 1; i < n_upper_rows; ++i) {
            const A* row = a_ptr + (ptrdiff_t)i * row_stride;  // left.row(i), stride col_stride
            A sum = dot<A>(i, row, col_stride, col, row_stride);
            A* ci = &col[(ptrdiff_t)i * row_stride];
            *ci = *ci - sum;
        }
        if (ul < nrows) {
            A* lower_col = col + (ptrdiff_t)ul * row_stride;
            gemv_notrans<A>(-ScalarTraits<A>::one(), a_ptr + (ptrdiff_t)ul * row_stride,
                            nrows - ul, ul, row_stride, col_stride, col, row_stride,
                            ScalarTraits<A>::one(), lower_col, row_stride);
            size_t max_row;
            R max_val;
            iamax<A>(nrows - ul, lower_col, row_stride, &max_row, &max_val);
            size_t pivot_row = ul + max_row;
            pivots[ul] = pivot_row;
            A pivot = col[(ptrdiff_t)pivot_row * row_stride];
            if (pivot == ScalarTraits<A>::zero()) {
                singular_row = (int64_t)ul;
            } else {
                A pivot_recip = ScalarTraits<A>::one() / pivot;
                if (pivot_row != ul) {
                    swap_rows(ul + 1, a_ptr + (ptrdiff_t)ul * row_stride,
                              a_ptr + (ptrdiff_t)pivot_row * row_stride, col_stride);
                }
                for (size_t row = ul + 1; row < nrows; ++row) {
                    A* e = &col[(ptrdiff_t)row * row_stride];
                    *e = *e * pivot_recip;
                }
            }
        }
    }
    return singular_row;
}

// ---- getrf dispatch (src/lapack/getrf.rs:12-27) -----------------------------
// ndarray `is_standard_layout` for Ix2: C-contiguous strides, with the usual
// exemption for axes of length <= 1 and for empty arrays.
inline bool is_standard_layout(size_t nrows, size_t ncols, ptrdiff_t rs, ptrdiff_t cs) {
    if (nrows == 0 || ncols == 0) return true;
    if (ncols != 1 && cs != 1) return false;
    if (nrows != 1 && rs != (ptrdiff_t)ncols) return false;
    return true;
}

template <class A>
inline int64_t getrf(A* a, size_t nrows, size_t ncols, ptrdiff_t rs, ptrdiff_t cs, size_t* pivots) {
    size_t dim_min = std::min(nrows, ncols);
    for (size_t i = 0; i < dim_min; ++i) pivots[i] = i;
    if (is_standard_layout(nrows, ncols, rs, cs)) return getrf_row_major<A>(a, nrows, ncols, rs, cs, pivots);
    return getrf_col_major<A>(a, nrows, ncols, rs, cs, pivots);
}

// ---- recursive_inner / getrf_recursive (src/lapack/getrf.rs:216-322, 30-40) --
// Test-only in the reference; the blocked order the CUDA path follows.  Returns
// 0 for Ok and (row) for Err(Singular(row)); note the reference's quirk that a
// singularity at row 0 is indistinguishable from Ok (`singular_row == 0`).
template <class A>
inline bool recursive_inner(A* a, size_t nrows, size_t ncols, ptrdiff_t rs, ptrdiff_t cs,
                            size_t* pivots, size_t* err_row) {
    using R = RealOf<A>;
    if (nrows == 0 || ncols == 0) return true;
    if (nrows == 1) {
        pivots[0] = 0;
        if (a[0] == ScalarTraits<A>::zero()) {
            *err_row = 0;
            return false;
        }
        return true;
    }
    if (ncols == 1) {
        size_t max_idx;
        R max_val;
        iamax<A>(nrows, a, rs, &max_idx, &max_val);
        pivots[0] = max_idx;
        if (max_val == R(0)) {
            *err_row = 0;
            return false;
        }
        if (max_idx != 0) std::swap(a[0], a[(ptrdiff_t)max_idx * rs]);
        if (max_val >= sfmin<R>()) {
            A alpha = ScalarTraits<A>::one() / a[0];
            for (size_t i = 1; i < nrows; ++i) a[(ptrdiff_t)i * rs] = a[(ptrdiff_t)i * rs] * alpha;
        } else {
            A pivot = a[0];
            for (size_t i = 1; i < nrows; ++i) a[(ptrdiff_t)i * rs] = a[(ptrdiff_t)i * rs] / pivot;
        }
        return true;
    }
    size_t left_cols = std::min(nrows, ncols) / 2;
    size_t right_cols = ncols - left_cols;
    size_t singular_row = 0;
    {
        size_t row = 0;
        if (!recursive_inner<A>(a, nrows, left_cols, rs, cs, pivots, &row)) singular_row = row;
    }
    laswp<A>(right_cols, a + cs * (ptrdiff_t)left_cols, rs, cs, 0, pivots, left_cols);
    trsm<A>(a, rs, cs, a + cs * (ptrdiff_t)left_cols, left_cols, right_cols, rs, cs);
    size_t min_dim = std::min(nrows, ncols);
    A* lower_left = a + rs * (ptrdiff_t)left_cols;
    A* upper_right = a + cs * (ptrdiff_t)left_cols;
    A* lower_right = lower_left + cs * (ptrdiff_t)left_cols;
    gemm<A>(-ScalarTraits<A>::one(), lower_left, nrows - left_cols, left_cols, rs, cs, upper_right,
            right_cols, rs, cs, lower_right, rs, cs);
    {
        size_t row = 0;
        if (!recursive_inner<A>(lower_right, nrows - left_cols, right_cols, rs, cs,
                                pivots + left_cols, &row)) {
            if (singular_row == 0) singular_row = left_cols + row;
        }
    }
    for (size_t i = left_cols; i < min_dim; ++i) pivots[i] += left_cols;
    laswp<A>(left_cols, a, rs, cs, left_cols, pivots, min_dim);
    if (singular_row == 0) return true;
    *err_row = singular_row;
    return false;
}

// ---- lapack::getrs (src/lapack/getrs.rs:12-38) ------------------------------
// x = b; laswp(x, p); row-oriented unit-lower forward sweep; row-oriented upper
// back sweep with a true divide by the diagonal.  Single right-hand side.
template <class A>
inline void getrs(const A* a, size_t n, ptrdiff_t rs, ptrdiff_t cs, const size_t* p, const A* b,
                  ptrdiff_t incb, A* x) {
    for (size_t i = 0; i < n; ++i) x[i] = b[(ptrdiff_t)i * incb];
    if (n > 0) laswp<A>(1, x, 1, 1, 0, p, n);
    for (size_t i = 0; i < n; ++i) {
        for (size_t k = 0; k < i; ++k) {
            A prod = a[(ptrdiff_t)i * rs + (ptrdiff_t)k * cs] * x[k];
            x[i] = x[i] - prod;
        }
    }
    for (size_t ii = n; ii-- > 0;) {
        for (size_t k = ii + 1; k < n; ++k) {
            A prod = a[(ptrdiff_t)ii * rs + (ptrdiff_t)k * cs] * x[k];
            x[ii] = x[ii] - prod;
        }
        x[ii] = x[ii] / a[(ptrdiff_t)ii * rs + (ptrdiff_t)ii * cs];
    }
}

// ---- lu::Factorized::into_pl (src/decomposition/lu.rs:107-153) --------------
// In-place P*L in the first min(m,n) columns, following the reference's
// cycle-walking loop verbatim (pivots is consumed).
template <class A>
inline void into_pl(A* lu, size_t nrows, size_t ncols, ptrdiff_t rs, ptrdiff_t cs,
                    std::vector<size_t> pivots) {
    if (pivots.size() < nrows) {
        for (size_t next = pivots.size(); next < nrows; ++next) pivots.push_back(next);
    }
    for (size_t i = pivots.size(); i-- > 0;) {
        size_t target = pivots[i];
        if (i == target) continue;
        pivots[i] = pivots[target];
        pivots[target] = i;
    }
    size_t pl_cols = std::min(nrows, ncols);
    if (pivots.empty()) return;
    auto at = [&](size_t r, size_t c) -> A& { return lu[(ptrdiff_t)r * rs + (ptrdiff_t)c * cs]; };
    size_t dst = 0, i = 0;
    const size_t done = pivots.size();
    for (;;) {
        size_t src = pivots[dst];
        for (size_t k = 0; k < std::min(src, pl_cols); ++k) at(dst, k) = at(src, k);
        if (src < pl_cols) at(dst, src) = ScalarTraits<A>::one();
        for (size_t k = src + 1; k < pl_cols; ++k) at(dst, k) = ScalarTraits<A>::zero();
        pivots[dst] = done;
        if (pivots[src] == done) {
            dst = i + 1;
            while (dst < done && pivots[dst] == done) ++dst;
            if (dst == done) break;
            i = dst;
        } else {
            dst = src;
        }
    }
}

// =============================================================================
// QR path: lapack::geqrf and decomposition::qr::Factorized.
// Pinned by the reference's own tests (tests/golden/lair_qr_golden.json, transcribed
// from src/lapack/larfg.rs:44-84, src/lapack/geqrf.rs:33-141,
// src/decomposition/qr.rs:92-199) in tests/test_oracle_qr.py.
// =============================================================================
template <class T> inline Cx<T> conj_of(Cx<T> a) { return {a.re, -a.im}; }
inline float conj_of(float a) { return a; }
inline double conj_of(double a) { return a; }
template <class T> inline Cx<T> times_real(Cx<T> a, T r) { return {a.re * r, a.im * r}; }   // Complex * T
inline float times_real(float a, float r) { return a * r; }
inline double times_real(double a, double r) { return a * r; }
template <class T> inline Cx<T> over_real(Cx<T> a, T r) { return {a.re / r, a.im / r}; }    // Complex / T
inline float over_real(float a, float r) { return a / r; }
inline double over_real(double a, double r) { return a / r; }
template <class T> inline Cx<T> from_real(T r, Cx<T>*) { return {r, T(0)}; }
inline float from_real(float r, float*) { return r; }
inline double from_real(double r, double*) { return r; }
template <class A> inline bool is_zero_s(A a) { return a == ScalarTraits<A>::zero(); }

// blas::nrm2 (src/blas/nrm2.rs:6-15): plain in-order sum of re^2 + im^2, then sqrt.
template <class A>
inline RealOf<A> nrm2(size_t n, const A* x, ptrdiff_t inc) {
    using R = RealOf<A>;
    R sum = R(0);
    for (size_t i = 0; i < n; ++i) {
        const A v = x[(ptrdiff_t)i * inc];
        sum = sum + (ScalarTraits<A>::re(v) * ScalarTraits<A>::re(v) + ScalarTraits<A>::im(v) * ScalarTraits<A>::im(v));
    }
    return std::sqrt(sum);
}
// lapack::lapy3 (src/lapack.rs:63-65)
template <class R> inline R lapy3(R x, R y, R z) { return std::sqrt(x * x + y * y + z * z); }

// lapack::larfg (src/lapack/larfg.rs:9-42): x is overwritten with v[1..], returns (beta, tau).
template <class A>
inline void larfg(A alpha, size_t n, A* x, ptrdiff_t inc, RealOf<A>* beta_out, A* tau_out) {
    using R = RealOf<A>;
    using S = ScalarTraits<A>;
    R x_norm = nrm2<A>(n, x, inc);
    if (x_norm == R(0) && S::im(alpha) == R(0)) {
        *beta_out = S::re(alpha);
        *tau_out = S::zero();
        return;
    }
    R beta = -std::copysign(lapy3<R>(S::re(alpha), S::im(alpha), x_norm), S::re(alpha));
    const R eps = std::numeric_limits<R>::epsilon() / (R(1) + R(1));   // Real::eps (src/scalar.rs:393-395)
    const R safe_min = sfmin<R>() / eps;
    int knt = 0;
    if (std::fabs(beta) < safe_min) {
        const R safe_min_recip = R(1) / safe_min;
        for (;;) {
            ++knt;
            for (size_t i = 0; i < n; ++i) x[(ptrdiff_t)i * inc] = times_real(x[(ptrdiff_t)i * inc], safe_min_recip);
            beta *= safe_min_recip;
            alpha = times_real(alpha, safe_min_recip);
            if (std::fabs(beta) >= safe_min || knt >= 20) break;
        }
        x_norm = nrm2<A>(n, x, inc);
        // literal: `alpha.square() + x_norm * x_norm` without a square root (larfg.rs:33)
        const R asq = S::re(alpha) * S::re(alpha) + S::im(alpha) * S::im(alpha);
        beta = -std::copysign(asq + x_norm * x_norm, S::re(alpha));
    }
    const A beta_a = from_real(beta, (A*)nullptr);
    const A tau = over_real(beta_a - alpha, beta);
    alpha = S::one() / (alpha - beta_a);
    for (size_t i = 0; i < n; ++i) x[(ptrdiff_t)i * inc] = x[(ptrdiff_t)i * inc] * alpha;
    beta *= std::pow(safe_min, (R)knt);
    *beta_out = beta;
    *tau_out = tau;
}

// lapack::larf::left (src/lapack/larf.rs:10-54): C := (I - tau v v^H) C, with the reference's
// trimming of v's trailing zeros and C's trailing zero columns (ilalc, src/lapack/ilal.rs:6-22).
template <class A>
inline void larf_left(size_t nv, const A* v, ptrdiff_t incv, A tau, A* c, size_t ncols, ptrdiff_t rs, ptrdiff_t cs) {
    using S = ScalarTraits<A>;
    if (is_zero_s(tau)) return;
    size_t last_v = nv;  // the reference unwraps: a zero vector panics; callers always pass v[0] = 1
    for (size_t i = nv; i-- > 0;)
        if (!is_zero_s(v[(ptrdiff_t)i * incv])) { last_v = i; break; }
    if (last_v == nv) return;
    auto at = [&](size_t r, size_t col) -> A& { return c[(ptrdiff_t)r * rs + (ptrdiff_t)col * cs]; };
    if (ncols == 0) return;
    size_t last_c = ncols;
    if (!is_zero_s(at(last_v, ncols - 1))) {
        last_c = ncols - 1;
    } else {
        for (size_t col = ncols; col-- > 0 && last_c == ncols;)
            for (size_t r = 0; r <= last_v; ++r)
                if (!is_zero_s(at(r, col))) { last_c = col; break; }
    }
    if (last_c == ncols) return;
    // w = C^H v (blas::gemv::conjtrans, src/blas/gemv.rs:58-88): fold from zero, then y = 0 + 1 * sum
    std::vector<A> w(last_c + 1);
    for (size_t j = 0; j <= last_c; ++j) {
        A sum = S::zero();
        for (size_t r = 0; r <= last_v; ++r) sum = sum + conj_of(at(r, j)) * v[(ptrdiff_t)r * incv];
        w[j] = S::zero() + S::one() * sum;
    }
    // C += (-tau) v w^H (blas::gerc, src/blas/gerc.rs:8-34)
    const A alpha = -tau;
    for (size_t i = 0; i <= last_v; ++i) {
        const A factor = alpha * v[(ptrdiff_t)i * incv];
        for (size_t j = 0; j <= last_c; ++j) at(i, j) = at(i, j) + factor * conj_of(w[j]);
    }
}

// lapack::geqrf (src/lapack/geqrf.rs:9-30): in place, tau has min(m, n) entries.
template <class A>
inline void geqrf(A* a, size_t m, size_t n, ptrdiff_t rs, ptrdiff_t cs, A* tau) {
    using S = ScalarTraits<A>;
    auto at = [&](size_t r, size_t c) -> A& { return a[(ptrdiff_t)r * rs + (ptrdiff_t)c * cs]; };
    const size_t min_dim = std::min(m, n);
    for (size_t i = 0; i < min_dim; ++i) {
        RealOf<A> beta;
        A t;
        larfg<A>(at(i, i), m - i - 1, &at(i, i) + rs, rs, &beta, &t);
        tau[i] = t;
        at(i, i) = S::one();
        std::vector<A> v(m - i);
        for (size_t r = i; r < m; ++r) v[r - i] = at(r, i);
        if (i + 1 < n) larf_left<A>(m - i, v.data(), 1, conj_of(t), &at(i, i + 1), n - i - 1, rs, cs);
        at(i, i) = from_real(beta, (A*)nullptr);
    }
}

// qr::Factorized::q (src/decomposition/qr.rs:27-59): q is m x m row-major, dense.
template <class A>
inline void qr_q(const A* qr, size_t m, size_t n, ptrdiff_t rs, ptrdiff_t cs, const A* tau, A* q) {
    using S = ScalarTraits<A>;
    const size_t k = std::min(m, n);
    auto Q = [&](size_t r, size_t c) -> A& { return q[r * m + c]; };
    for (size_t j = 0; j < k; ++j)
        for (size_t i = 0; i < m; ++i) Q(i, j) = qr[(ptrdiff_t)i * rs + (ptrdiff_t)j * cs];
    for (size_t j = k; j < m; ++j) {
        for (size_t l = 0; l < m; ++l) Q(l, j) = S::zero();
        Q(j, j) = S::one();
    }
    for (size_t i = k; i-- > 0;) {
        if (i + 1 < m) {
            Q(i, i) = S::one();
            std::vector<A> v(m - i);
            for (size_t r = i; r < m; ++r) v[r - i] = Q(r, i);
            larf_left<A>(m - i, v.data(), 1, tau[i], &Q(i, i + 1), m - i - 1, (ptrdiff_t)m, 1);
            const A nt = -tau[i];
            for (size_t r = i + 1; r < m; ++r) Q(r, i) = Q(r, i) * nt;
        }
        Q(i, i) = S::one() - tau[i];
        for (size_t l = 0; l < i; ++l) Q(l, i) = S::zero();
    }
}

}  // namespace lair_oracle
