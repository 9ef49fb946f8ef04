"""GPU: Householder QR (lapack::geqrf, qr::Factorized) through the C ABI against the reference's own
golden vectors and the CPU oracle.

Bars: the reference's test epsilons on its golden cases; |QR - QR_oracle| and |tau - tau_oracle| to rounding
(the CUDA path sums nrm2 and the column dot products in parallel order); scaled factorization error
||A - QR||_F / (max(m, n) eps ||A||_F) and orthogonality ||Q^H Q - I||_F / (m eps) <= 10x the oracle's own.
"""
import numpy as np
import pytest

import oracle
from qr_golden_util import check_geqrf_case, check_qr_case, load_qr_golden, mat, qr_errors

pytestmark = pytest.mark.gpu
G = load_qr_golden()


@pytest.fixture(scope="module")
def lair():
    import lair_b200
    return lair_b200


def _rand(rng, shape, dt, dist="uniform"):
    a = rng.uniform(0, 10, size=shape) if dist == "uniform" else rng.standard_normal(shape)
    if np.issubdtype(dt, np.complexfloating):
        a = a + 1j * (rng.uniform(0, 10, size=shape) if dist == "uniform" else rng.standard_normal(shape))
    return a.astype(dt)


@pytest.mark.parametrize("case", G["geqrf"], ids=lambda c: c["name"])
def test_geqrf_golden(lair, case):
    a = mat(case, "a")
    for make in (lambda x: x, np.asfortranarray, lambda x: np.ascontiguousarray(x[::-1, ::-1])[::-1, ::-1]):
        qr = make(a.copy())
        tau = lair.lapack.geqrf(qr)
        check_geqrf_case(case, qr, tau)


@pytest.mark.parametrize("case", G["qr"], ids=lambda c: c["name"])
def test_qr_factorized_golden(lair, case):
    a = mat(case, "a")
    f = lair.decomposition.qr.Factorized.from_(a)
    check_geqrf_case(case, f.qr, f.tau)
    check_qr_case(case, f.q(), f.r())
    assert np.array_equal(a, mat(case, "a"))


@pytest.mark.parametrize("dt", [np.float64, np.float32, np.complex128, np.complex64])
@pytest.mark.parametrize("shape", [(1, 1), (2, 2), (33, 33), (100, 100), (300, 300), (500, 120), (120, 500), (1000, 64), (257, 129)])
def test_geqrf_matches_oracle(lair, dt, shape):
    rng = np.random.default_rng(shape[0] * 11 + shape[1])
    a0 = _rand(rng, shape, dt)
    qr = a0.copy()
    tau = lair.lapack.geqrf(qr)
    ref = a0.copy()
    tau_o = oracle.geqrf(ref)
    eps = np.finfo(dt).eps
    scale = np.max(np.abs(ref))
    tol = 50 * eps * max(shape) * scale
    assert tau.shape == tau_o.shape
    assert np.max(np.abs(qr - ref)) <= tol, np.max(np.abs(qr - ref)) / scale
    assert np.max(np.abs(tau - tau_o)) <= 50 * eps * max(shape)
    q, r = lair.lapack.qr_q(qr, tau), np.triu(qr)
    q_o = oracle.qr_q(ref, tau_o)
    assert q.shape == (shape[0], shape[0])
    assert np.max(np.abs(q - q_o)) <= 50 * eps * max(shape)
    fact, orth = qr_errors(a0, q, r)
    fact_o, orth_o = qr_errors(a0, q_o, np.triu(ref))
    assert fact <= 10 * max(fact_o, 0.01), (fact, fact_o)
    assert orth <= 10 * max(orth_o, 0.01), (orth, orth_o)


def test_geqrf_edges(lair):
    # empty shapes: tau is empty, nothing is touched (geqrf.rs:14-15)
    for shape in ((0, 0), (0, 4), (4, 0)):
        a = np.zeros(shape)
        assert lair.lapack.geqrf(a).shape == (0,)
    # a zero column below the diagonal with a real diagonal: H = I, tau = 0 (larfg.rs:15-17)
    a = np.array([[3.0, 1.0], [0.0, 2.0]])
    qr = a.copy()
    tau = lair.lapack.geqrf(qr)
    assert np.array_equal(qr, a) and np.array_equal(tau, np.zeros(2))
    # zero matrix: every reflector is the identity; Q = I
    z = np.zeros((5, 3), dtype=np.float32)
    tau = lair.lapack.geqrf(z)
    assert not z.any() and not tau.any()
    assert np.array_equal(lair.lapack.qr_q(z, tau), np.eye(5, dtype=np.float32))
    # tiny entries: the safe-minimum rescaling loop of larfg (larfg.rs:22-34), restated literally like the oracle
    t0 = np.array([[1e-300, 2e-300], [3e-300, 4e-300]])
    t, t_ref = t0.copy(), t0.copy()
    tau = lair.lapack.geqrf(t)
    tau_o = oracle.geqrf(t_ref)
    assert np.allclose(t, t_ref, rtol=1e-12, atol=0) and np.allclose(tau, tau_o, rtol=1e-12, atol=0)
    # non-square Factorized: shapes of q and r (qr.rs:135-199)
    f = lair.decomposition.qr.Factorized.from_(np.arange(12, dtype=np.float64).reshape(3, 4) + np.eye(3, 4))
    assert f.q().shape == (3, 3) and f.r().shape == (3, 4) and f.tau.shape == (3,)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(64, 64), (200, 200), (1000, 1000), (3000, 200), (200, 900), (2049, 97), (777, 333)])
def test_geqrf_blocked_equals_unblocked(lair, dt, shape):
    """Compact-WY blocks (cluster panel + three GEMMs per block, qr_blocked.cu) against the one-reflector-at-a-time
    loop (qr.cu) and the oracle: R, the reflectors and tau to rounding; factorization error and orthogonality within
    10x the oracle's."""
    from lair_b200 import _ffi
    rng = np.random.default_rng(shape[0] + 13 * shape[1])
    a0 = _rand(rng, shape, dt, "normal")
    out = {}
    default = _ffi.get_option("qr_blocked")
    try:
        for v in (0, 1):
            _ffi.set_option("qr_blocked", v)
            qr = a0.copy()
            out[v] = (qr, lair.lapack.geqrf(qr))
        _ffi.set_option("lookahead", 0)  # the same blocks without the panel / update overlap: identical bits
        qr = a0.copy()
        tau = lair.lapack.geqrf(qr)
        assert np.array_equal(qr, out[1][0]) and np.array_equal(tau, out[1][1])
    finally:
        _ffi.set_option("qr_blocked", default)
        _ffi.set_option("lookahead", 1)
    eps = np.finfo(dt).eps
    scale = np.max(np.abs(out[0][0]))
    assert np.max(np.abs(out[0][0] - out[1][0])) <= 100 * eps * max(shape) * scale
    assert np.max(np.abs(out[0][1] - out[1][1])) <= 100 * eps * max(shape)
    if shape[0] <= 1000:
        ref = a0.copy()
        tau_o = oracle.geqrf(ref)
        assert np.max(np.abs(out[1][0] - ref)) <= 100 * eps * max(shape) * scale
        q = lair.lapack.qr_q(out[1][0], out[1][1])
        fact, orth = qr_errors(a0, q, np.triu(out[1][0]))
        fact_o, orth_o = qr_errors(a0, oracle.qr_q(ref, tau_o), np.triu(ref))
        assert fact <= 10 * max(fact_o, 0.01) and orth <= 10 * max(orth_o, 0.01), (fact, fact_o, orth, orth_o)


@pytest.mark.parametrize("dt,shape", [(np.float64, (14000, 96)), (np.float32, (30000, 64)), (np.complex128, (9000, 40))])
def test_geqrf_tall_panels(lair, dt, shape):
    """Panels taller than one cluster's shared memory (f64 > ~12 800 rows, f32 > ~25 600): the blocked sweep factors the
    panel with the one-reflector loop, whose reflector application splits the rows over CTAs (two passes with in-order
    partial sums), and rebuilds T from tau and V^T V.  R, the reflectors and tau against the oracle; A = QR checked
    through R^H R = A^H A (Q would be m x m)."""
    rng = np.random.default_rng(shape[0])
    a0 = _rand(rng, shape, dt, "normal")
    qr = a0.copy()
    tau = lair.lapack.geqrf(qr)
    ref = a0.copy()
    tau_o = oracle.geqrf(ref)
    eps = np.finfo(dt).eps
    scale = np.max(np.abs(ref))
    assert np.max(np.abs(qr - ref)) <= 100 * eps * np.sqrt(shape[0]) * shape[1] * scale
    assert np.max(np.abs(tau - tau_o)) <= 100 * eps * np.sqrt(shape[0]) * shape[1]
    wide = np.complex128 if np.iscomplexobj(a0) else np.float64
    r = np.triu(qr[: shape[1]]).astype(wide)
    a = a0.astype(wide)
    gram = np.linalg.norm(r.conj().T @ r - a.conj().T @ a) / (np.linalg.norm(a) ** 2)
    assert gram <= 100 * eps, gram
