"""Helpers shared by the golden-vector tests (oracle on CPU, CUDA path on GPU)."""
import json
import os

import numpy as np

_DT = {"f32": np.float32, "f64": np.float64, "c64": np.complex64, "c128": np.complex128}
GOLDEN_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lair_golden.json")


def load_golden():
    with open(GOLDEN_PATH) as f:
        return json.load(f)


def dtype_of(case):
    return _DT[case["dtype"]]


def build_input(case):
    """Materialise a case's input with the memory layout the reference test used.

    layout "row": C-order as written.  "swap_axes": the literal is stored C-order and then
    axes are swapped (src/lapack/getrf.rs:440), so the logical matrix is its transpose with
    strides (1, n).  "invert_both": both axes reversed (src/lapack/getrf.rs:466-467).
    """
    dt = dtype_of(case)
    if "a_re" in case:
        a = np.array(case["a_re"], dtype=dt) + 1j * np.array(case["a_im"], dtype=dt)
        a = a.astype(dt)
    elif "shape" in case:
        a = np.zeros(case["shape"], dtype=dt)
    else:
        a = np.array(case["a"], dtype=dt)
    layout = case.get("layout", "row")
    if layout == "swap_axes":
        a = a.swapaxes(0, 1)
    elif layout == "invert_both":
        a = a[::-1, ::-1]
    elif layout != "row":
        raise ValueError(layout)
    return a


def expected_lu(case):
    dt = dtype_of(case)
    if "lu_re" in case:
        return (np.array(case["lu_re"], dtype=dt) + 1j * np.array(case["lu_im"], dtype=dt)).astype(dt), case.get("abs_eps", 0.0)
    if "lu_exact" in case:
        return np.array(case["lu_exact"], dtype=dt), 0.0
    if "lu" in case:
        return np.array(case["lu"], dtype=dt), case["abs_eps"]
    return None, None


def check_getrf_case(case, pivots, singular, a_after):
    if "pivots" in case:
        assert list(pivots) == case["pivots"], (case["name"], pivots)
    if "singular" in case:
        assert singular == case["singular"], (case["name"], singular)
    exp, eps = expected_lu(case)
    if exp is not None:
        if eps == 0.0:
            assert np.array_equal(np.asarray(a_after), exp), (case["name"], a_after)
        else:
            assert np.max(np.abs(np.asarray(a_after) - exp)) <= eps, (case["name"], a_after)


# lu::Factorized::{p,l,u} restated on the host for checking (src/decomposition/lu.rs:28-72).
def lu_p(nrows, pivots, dtype):
    perm = list(range(nrows))
    for i, p in enumerate(pivots):
        if i != p:
            perm[i], perm[p] = perm[p], perm[i]
    out = np.zeros((nrows, nrows), dtype=dtype)
    for i, pv in enumerate(perm):
        out[pv, i] = 1
    return out


def lu_l(lu):
    m, n = lu.shape
    rank = min(m, n)
    out = np.zeros((m, rank), dtype=lu.dtype)
    for i in range(m):
        out[i, : min(i, rank)] = lu[i, : min(i, rank)]
        if i < rank:
            out[i, i] = 1
    return out


def lu_u(lu):
    m, n = lu.shape
    rank = min(m, n)
    out = np.zeros((rank, n), dtype=lu.dtype)
    for i in range(rank):
        out[i, i:] = lu[i, i:]
    return out


def apply_pivots_rows(a, pivots):
    """P*A for sequential interchanges (what laswp does to rows)."""
    a = np.array(a, copy=True)
    for i, p in enumerate(pivots):
        if i != p:
            a[[i, p]] = a[[p, i]]
    return a


def backward_error(a0, lu, pivots):
    """Scaled backward error ||P A - L U||_F / (n * eps * ||A||_F), eps = Real::eps (2^-53 / 2^-24)."""
    m, n = a0.shape
    real = np.finfo(a0.dtype).dtype
    eps = np.finfo(real).eps / 2
    pa = apply_pivots_rows(a0, pivots).astype(np.complex128 if np.iscomplexobj(a0) else np.float64)
    l = lu_l(lu).astype(pa.dtype)
    u = lu_u(lu).astype(pa.dtype)
    num = np.linalg.norm(pa - l @ u)
    den = max(m, n) * eps * np.linalg.norm(pa)
    return float(num / den) if den > 0 else 0.0
