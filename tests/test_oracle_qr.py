"""CPU: the QR part of the oracle against every golden vector the reference's own tests hold for
geqrf / larfg / qr::Factorized (tests/golden/lair_qr_golden.json), plus properties on random inputs."""
import numpy as np
import pytest

import oracle
from qr_golden_util import check_geqrf_case, check_qr_case, cvec, load_qr_golden, mat, qr_errors

G = load_qr_golden()


@pytest.mark.parametrize("case", G["larfg"], ids=lambda c: c["name"])
def test_oracle_larfg_golden(case):
    x = cvec(case, "x")
    alpha = complex(*case["alpha"]) if np.iscomplexobj(x) else case["alpha"][0]
    beta, tau = oracle.larfg(alpha, x)
    eps = case["eps"]
    assert abs(beta - case["beta"]) <= max(eps, 0.0)
    assert abs(complex(tau) - complex(*case["tau"])) <= max(eps, 0.0)
    assert x.shape[0] == len(case["x_out"])
    if x.size:
        assert np.max(np.abs(x - cvec(case, "x_out"))) <= eps


@pytest.mark.parametrize("case", G["geqrf"], ids=lambda c: c["name"])
def test_oracle_geqrf_golden(case):
    a = mat(case, "a")
    qr = a.copy()
    tau = oracle.geqrf(qr)
    check_geqrf_case(case, qr, tau)
    # the same through a column-major and a reversed view (any strides)
    for make in (np.asfortranarray, lambda x: np.ascontiguousarray(x[::-1, ::-1])[::-1, ::-1]):
        v = make(a.copy())
        t = oracle.geqrf(v)
        assert np.array_equal(v, qr) and np.array_equal(t, tau)


@pytest.mark.parametrize("case", G["qr"], ids=lambda c: c["name"])
def test_oracle_qr_factorized_golden(case):
    a = mat(case, "a")
    qr = a.copy()
    tau = oracle.geqrf(qr)
    check_geqrf_case(case, qr, tau)
    check_qr_case(case, oracle.qr_q(qr, tau), oracle.qr_r(qr))


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.complex64, np.complex128])
@pytest.mark.parametrize("shape", [(1, 1), (6, 6), (9, 4), (4, 9), (40, 25), (0, 3)])
def test_oracle_qr_properties(dt, shape):
    rng = np.random.default_rng(shape[0] * 17 + shape[1])
    a0 = rng.uniform(0, 10, size=shape)
    if np.issubdtype(dt, np.complexfloating):
        a0 = a0 + 1j * rng.uniform(0, 10, size=shape)
    a0 = a0.astype(dt)
    qr = a0.copy()
    tau = oracle.geqrf(qr)
    assert tau.shape == (min(shape),)
    if 0 in shape:
        return
    q, r = oracle.qr_q(qr, tau), oracle.qr_r(qr)
    assert q.shape == (shape[0], shape[0]) and r.shape == shape
    fact, orth = qr_errors(a0, q, r)
    assert fact < 5 and orth < 5, (fact, orth)
    assert np.allclose(np.diag(r[: min(shape), : min(shape)]).imag, 0)  # beta is real (larfg.rs:18)
