"""GPU: the reference's golden vectors through the C ABI (same cases as test_oracle_golden.py)."""
import numpy as np
import pytest

import oracle
from golden_util import build_input, check_getrf_case, dtype_of, load_golden
from test_oracle_golden import _same_layout_copy

pytestmark = pytest.mark.gpu
G = load_golden()


@pytest.fixture(scope="module")
def lair():
    import lair_b200
    return lair_b200


@pytest.mark.parametrize("case", G["getrf"], ids=lambda c: c["name"])
def test_getrf_golden(lair, case):
    a = build_input(case)
    a = a.copy(order="K") if case.get("layout", "row") == "row" else _same_layout_copy(a)
    piv, sing = lair.lapack.getrf(a)
    check_getrf_case(case, piv, sing, a)
    # and identical to the oracle on the same input
    b = build_input(case)
    b = b.copy(order="K") if case.get("layout", "row") == "row" else _same_layout_copy(b)
    piv_o, sing_o = oracle.getrf(b)
    assert piv == piv_o and sing == sing_o
    if case.get("layout", "row") == "row":
        assert np.array_equal(a, b)  # standard layout: bit-identical to the reference's row-major body


@pytest.mark.parametrize("case", G["getrs"], ids=lambda c: c["name"])
def test_getrs_golden(lair, case):
    a = np.array(case["a"], dtype=dtype_of(case))
    piv, _ = lair.lapack.getrf(a)
    if "pivots" in case:
        assert piv == case["pivots"]
    b = np.array(case["b"], dtype=a.dtype)
    x = lair.lapack.getrs(a, piv, b)
    exp = np.array(case["x"])
    if "max_relative" in case:
        assert np.all(np.abs(x - exp) <= case["max_relative"] * np.maximum(np.abs(x), np.abs(exp)))
    else:
        assert np.all(np.abs(x - exp) <= case["abs_eps"])
    assert np.array_equal(x, oracle.getrs(a, piv, b))  # small systems: same operation order as getrs.rs


@pytest.mark.parametrize("case", G["lu"], ids=lambda c: c["name"])
def test_lu_factorized_golden(lair, case):
    a = np.array(case["a"], dtype=dtype_of(case))
    f = lair.decomposition.lu.Factorized.from_(a)
    p, l, u = f.p(), f.l(), f.u()
    assert list(p.shape) == case["p_shape"] and list(l.shape) == case["l_shape"] and list(u.shape) == case["u_shape"]
    for r, c in case["p_ones"]:
        assert p[r, c] == 1.0
    for name, mat in (("l", l), ("u", u)):
        for r, c, v in case[name]:
            assert abs(mat[r, c] - v) <= case["max_relative"] * max(abs(mat[r, c]), abs(v)), (name, r, c)
    assert not f.is_singular()


@pytest.mark.parametrize("case", G["into_pl"], ids=lambda c: c["name"])
def test_into_pl_golden(lair, case):
    a = np.array(case["a"], dtype=dtype_of(case))
    pl = lair.decomposition.lu.Factorized.from_(a).into_pl()
    k = min(a.shape)
    assert np.array_equal(pl[:, :k], np.array(case["pl"], dtype=a.dtype))


def test_solve_doctest(lair):
    c = G["solve"][0]
    x = lair.equation.solve(np.array(c["a"], dtype=np.float64), np.array(c["b"], dtype=np.float64))
    assert np.array_equal(x, np.array(c["x_exact"]))


def test_solve_errors(lair):
    from lair_b200 import InvalidInput
    with pytest.raises(InvalidInput.Shape):
        lair.equation.solve(np.zeros((2, 3)), np.zeros(2))
    with pytest.raises(InvalidInput.Shape):
        lair.equation.solve(np.eye(3), np.zeros(2))
    with pytest.raises(InvalidInput.Value):
        lair.equation.solve(np.ones((2, 2)), np.ones(2))
    f = lair.decomposition.lu.Factorized.from_(np.eye(3))
    with pytest.raises(InvalidInput.Shape):
        f.solve(np.zeros(4))
    with pytest.raises(AssertionError):
        lair.lapack.getrs(np.eye(3), [0, 1], np.zeros(3))
