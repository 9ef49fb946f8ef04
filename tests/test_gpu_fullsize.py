"""GPU: BASELINE.json's full sizes through size-independent properties.

The oracle needs minutes-to-hours at these sizes, so parity is carried by (a) the scaled
backward error ||PA - LU|| / (n eps ||A||) and the solve residual, evaluated on the device
in f64, against the bound 10x the oracle's own value at the largest common size (0.03 at
n=256 -> bound 0.5), (b) pivots identical to LAPACK's dgetrf (same first-max rule on
continuous data) where the host can afford it, and (c) encode -> decode round trips:
solve(A, A @ x) == x.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BOUND = 0.5


def _device_backward_error(a0, lu, piv):
    import torch
    n = a0.shape[0]
    dt = torch.float64
    A = torch.from_numpy(a0).cuda().to(dt)
    LU = torch.from_numpy(lu).cuda().to(dt)
    perm = np.arange(n)
    for i, p in enumerate(piv):
        if i != p:
            perm[i], perm[p] = perm[p], perm[i]
    PA = A[torch.from_numpy(perm).cuda()]
    L = torch.tril(LU, -1) + torch.eye(n, dtype=dt, device="cuda")
    U = torch.triu(LU)
    num = torch.linalg.norm(PA - L @ U)
    eps = np.finfo(a0.dtype).eps / 2
    return float(num / (n * eps * torch.linalg.norm(PA)))


def test_c2_n8192_f64_getrf_getrs():
    import lair_b200
    rng = np.random.default_rng(1)
    n, nrhs = 8192, 64
    a0 = rng.uniform(0, 10, size=(n, n))
    xs = rng.uniform(0, 10, size=(n, nrhs))
    b = a0 @ xs
    lu = a0.copy()
    piv, sing = lair_b200.lapack.getrf(lu)
    assert sing is None
    assert _device_backward_error(a0, lu, piv) <= BOUND
    x = lair_b200.lapack.getrs(lu, piv, b)
    eps = np.finfo(np.float64).eps / 2
    res = np.linalg.norm(a0 @ x - b) / (np.linalg.norm(a0) * np.linalg.norm(x) * n * eps)
    assert res <= BOUND
    scipy_linalg = pytest.importorskip("scipy.linalg")
    _, piv_l = scipy_linalg.lu_factor(a0, check_finite=False)
    assert piv == [int(p) for p in piv_l]


def test_c5b_n4096_f32_backward_error():
    import lair_b200
    rng = np.random.default_rng(6)
    n = 4096
    a0 = rng.uniform(0, 10, size=(n, n)).astype(np.float32)
    lu = a0.copy()
    piv, sing = lair_b200.lapack.getrf(lu)
    assert sing is None
    assert _device_backward_error(a0, lu, piv) <= BOUND


def test_c3_batched_1e6_roundtrip():
    """10^6 x (32x32): checksum parity against the oracle on a strided sample + P A = L U everywhere."""
    import lair_b200
    import oracle
    import torch
    rng = np.random.default_rng(3)
    batch = 1_000_000
    a0 = rng.uniform(0, 10, size=(batch, 32, 32)).astype(np.float32)
    a = a0.copy()
    ipiv, info = lair_b200.lapack.getrf_batched(a)
    assert np.all(info == -1)
    idx = np.arange(0, batch, 97)
    ref = a0[idx].copy()
    piv_o, _ = oracle.getrf_batched(ref)
    assert np.array_equal(ipiv[idx], piv_o.astype(np.int32))
    assert np.array_equal(a[idx], ref)
    # reconstruction on the device in f64 for every matrix
    LU = torch.from_numpy(a).cuda().double()
    L = torch.tril(LU, -1) + torch.eye(32, dtype=torch.float64, device="cuda")
    U = torch.triu(LU)
    rec = L @ U
    perm = np.tile(np.arange(32), (batch, 1))
    for j in range(32):
        p = ipiv[:, j]
        rows = np.arange(batch)
        tmp = perm[rows, j].copy()
        perm[rows, j] = perm[rows, p]
        perm[rows, p] = tmp
    PA = torch.gather(torch.from_numpy(a0).cuda().double(), 1, torch.from_numpy(perm).cuda()[:, :, None].expand(-1, -1, 32))
    err = torch.linalg.norm((PA - rec).reshape(batch, -1), dim=1) / torch.linalg.norm(PA.reshape(batch, -1), dim=1)
    assert float(err.max()) <= 32 * (np.finfo(np.float32).eps / 2) * 10
