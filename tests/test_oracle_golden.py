"""CPU: pin the oracle against every golden vector the reference's own tests hold."""
import numpy as np
import pytest

import oracle
from golden_util import (apply_pivots_rows, backward_error, build_input, check_getrf_case, dtype_of,
                         load_golden, lu_l, lu_p, lu_u)

G = load_golden()


@pytest.mark.parametrize("case", G["getrf"], ids=lambda c: c["name"])
def test_getrf_golden(case):
    a = build_input(case)
    a = a.copy(order="K") if case.get("layout", "row") == "row" else _same_layout_copy(a)
    piv, sing = oracle.getrf(a)
    check_getrf_case(case, piv, sing, a)


def _same_layout_copy(view):
    """Copy a strided view into fresh memory with the SAME stride signs/order (so the
    reference would still take its non-standard-layout branch)."""
    base = np.array(view, copy=True)  # C-order logical copy
    if view.strides[1] > view.strides[0] > 0:  # transposed
        buf = np.ascontiguousarray(base.T)
        out = buf.T
    elif view.strides[0] < 0 and view.strides[1] < 0:
        buf = np.ascontiguousarray(base[::-1, ::-1])
        out = buf[::-1, ::-1]
    else:
        raise AssertionError(view.strides)
    assert np.array_equal(out, view) and out.strides == view.strides
    return out


@pytest.mark.parametrize("case", G["getrf"], ids=lambda c: c["name"])
@pytest.mark.parametrize("variant", ["row", "col"])
def test_getrf_both_variants_agree_with_golden(case, variant):
    """Both live bodies (row-major right-looking, left-looking) reproduce every golden case."""
    a = np.array(build_input(case), copy=True)
    piv, sing = oracle.getrf(a, variant=variant)
    check_getrf_case(case, piv, sing, a)


@pytest.mark.parametrize("case", G["getrf_recursive"], ids=lambda c: c["name"])
def test_getrf_recursive_golden(case):
    a = np.array(build_input(case), copy=True)
    piv, err = oracle.getrf_recursive(a)
    assert err is None
    assert piv == case["pivots"]
    check_getrf_case({k: v for k, v in case.items() if k != "pivots"}, piv, None, a)


def test_iamax_golden():
    c = G["iamax"][0]
    idx, mx = oracle.iamax(np.array(c["x"], dtype=np.float64))
    assert (idx, mx) == (c["idx"], c["max"])


@pytest.mark.parametrize("case", G["getrs"], ids=lambda c: c["name"])
def test_getrs_golden(case):
    a = np.array(case["a"], dtype=dtype_of(case))
    piv, _ = oracle.getrf(a)
    if "pivots" in case:
        assert piv == case["pivots"]
    x = oracle.getrs(a, piv, np.array(case["b"], dtype=a.dtype))
    exp = np.array(case["x"])
    if "max_relative" in case:
        assert np.all(np.abs(x - exp) <= case["max_relative"] * np.maximum(np.abs(x), np.abs(exp)))
    else:
        assert np.all(np.abs(x - exp) <= case["abs_eps"])


@pytest.mark.parametrize("case", G["lu"], ids=lambda c: c["name"])
def test_lu_factorized_golden(case):
    a = np.array(case["a"], dtype=dtype_of(case))
    piv, _ = oracle.getrf(a)
    p, l, u = lu_p(a.shape[0], piv, a.dtype), lu_l(a), lu_u(a)
    assert list(p.shape) == case["p_shape"] and list(l.shape) == case["l_shape"] and list(u.shape) == case["u_shape"]
    for r, c in case["p_ones"]:
        assert p[r, c] == 1.0
    for name, mat in (("l", l), ("u", u)):
        for r, c, v in case[name]:
            assert abs(mat[r, c] - v) <= case["max_relative"] * max(abs(mat[r, c]), abs(v)), (name, r, c, mat[r, c])


@pytest.mark.parametrize("case", G["into_pl"], ids=lambda c: c["name"])
def test_into_pl_golden(case):
    a = np.array(case["a"], dtype=dtype_of(case))
    piv, _ = oracle.getrf(a)
    oracle.into_pl(a, piv)
    k = min(a.shape)
    assert np.array_equal(a[:, :k], np.array(case["pl"], dtype=a.dtype))


def test_solve_doctest():
    c = G["solve"][0]
    a = np.array(c["a"], dtype=np.float64)
    piv, sing = oracle.getrf(a)
    assert sing is None
    x = oracle.getrs(a, piv, np.array(c["b"], dtype=np.float64))
    assert np.array_equal(x, np.array(c["x_exact"]))


# ---- semantics the reference has but its tests do not pin -------------
def test_singular_is_last_zero_pivot_step():
    a = np.zeros((3, 3))
    piv, sing = oracle.getrf(a)
    assert piv == [0, 1, 2] and sing == 2
    a = np.array([[1.0, 2, 3], [2, 4, 6], [3, 6, 9]])  # rank 1: zero pivots at steps 1 and 2
    piv, sing = oracle.getrf(a)
    assert sing == 2


def test_first_max_tie_break_and_nan():
    assert oracle.iamax(np.array([2.0, -2.0, 2.0]))[0] == 0
    assert oracle.iamax(np.array([np.nan, 1.0, np.nan]))[0] == 1
    assert oracle.iamax(np.array([np.nan, np.nan])) == (0, 0.0)
    assert oracle.iamax(np.array([1 + 1j, -2 + 0j, 0 - 2j], dtype=np.complex128)) == (0, 2.0)


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.complex64, np.complex128])
@pytest.mark.parametrize("shape", [(1, 1), (7, 7), (33, 33), (100, 100), (40, 17), (17, 40)])
def test_layouts_agree_and_backward_error_small(dt, shape):
    rng = np.random.default_rng(1234)
    a0 = rng.uniform(0, 10, size=shape)
    if np.issubdtype(dt, np.complexfloating):
        a0 = a0 + 1j * rng.uniform(0, 10, size=shape)
    a0 = a0.astype(dt)
    a_row = a0.copy()
    piv_r, sing_r = oracle.getrf(a_row)
    a_col = np.asfortranarray(a0)
    assert a_col.shape == (1, 1) or not a_col.flags.c_contiguous or min(shape) == 1
    piv_c, sing_c = oracle.getrf(a_col)
    a_rec = a0.copy()
    piv_x, err = oracle.getrf_recursive(a_rec)
    assert sing_r is None and sing_c is None
    assert piv_r == piv_c == piv_x[: min(shape)]
    tol = 200 * np.finfo(dt).eps * 10 * max(shape)
    assert np.max(np.abs(a_row - a_col)) <= tol
    assert np.max(np.abs(a_row - a_rec)) <= tol
    assert backward_error(a0, a_row, piv_r) < 1.0
    assert backward_error(a0, a_col, piv_c) < 1.0


def test_getrs_residual_random():
    rng = np.random.default_rng(7)
    n = 200
    a0 = rng.uniform(0, 10, size=(n, n))
    b = rng.uniform(0, 10, size=n)
    a = a0.copy()
    piv, sing = oracle.getrf(a)
    assert sing is None
    x = oracle.getrs(a, piv, b)
    r = np.linalg.norm(a0 @ x - b) / (np.linalg.norm(a0) * np.linalg.norm(x) * n * np.finfo(np.float64).eps)
    assert r < 1.0
    # strided b and a
    x2 = oracle.getrs(np.asfortranarray(a), piv, np.repeat(b, 2)[::2])
    assert np.array_equal(x, x2)


def test_laswp_matches_sequential_swaps():
    rng = np.random.default_rng(3)
    a = rng.standard_normal((9, 4))
    piv = [3, 1, 8, 3, 4]
    exp = apply_pivots_rows(a, piv)
    oracle.laswp(a, piv)
    assert np.array_equal(a, exp)


def test_pivots_agree_with_lapack_dgetrf():
    """Independent cross-check: LAPACK's dgetrf picks the same first-max pivots on continuous data."""
    scipy_linalg = pytest.importorskip("scipy.linalg")
    rng = np.random.default_rng(0)
    a0 = rng.uniform(0, 10, size=(100, 100))  # benches/getrf.rs:11-12 shape and distribution
    a = a0.copy()
    piv, sing = oracle.getrf(a)
    lu, piv_l = scipy_linalg.lu_factor(a0)
    assert sing is None and piv == [int(p) for p in piv_l]
    assert np.max(np.abs(lu - a)) < 1e-10
