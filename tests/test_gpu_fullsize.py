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


def _getrf_dev(a_t):
    """lair_b200_{s,d}getrf_dev on a torch CUDA matrix (in place); returns (ipiv tensor, info)."""
    import torch
    from lair_b200 import _ffi
    m, n = a_t.shape
    pfx = "d" if a_t.dtype == torch.float64 else "s"
    ipiv = torch.empty(min(m, n), dtype=torch.int32, device="cuda")
    info = torch.empty(1, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    _ffi.check(getattr(_ffi.lib(), f"lair_b200_{pfx}getrf_dev")(m, n, a_t.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), stream))
    torch.cuda.synchronize()
    _ffi.check_fault(stream)
    return ipiv, int(info.item())


def test_c5a_tall_skinny_262144x1024_f32():
    """BASELINE configs[4], panel-dominated shape (the reference pins tall shapes at getrf.rs:505-524): device-side
    backward error <= 0.5 (10x the oracle's own at the largest common size), pivots in range, and -- on the leading
    2048-row block of the SAME columns, where the oracle finishes in seconds -- the tall-panel code path against the oracle."""
    import torch
    import devcheck
    import lair_b200
    import oracle
    m, n = 262144, 1024
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5)
    a0 = torch.rand(m, n, dtype=torch.float32, device="cuda", generator=gen) * 10
    a = a0.clone()
    ipiv, info = _getrf_dev(a)
    assert info == -1
    piv = ipiv.cpu().numpy()
    assert (piv >= np.arange(n)).all() and (piv < m).all()
    assert devcheck.backward_error_dev(a0, a, piv) <= BOUND
    # a tall problem the oracle can do: 16384 x 64 through the same tall-panel kernels
    ms, ns = 16384, 64
    h0 = a0[:ms, :ns].cpu().numpy().copy()
    h = h0.copy()
    piv_s, sing = lair_b200.lapack.getrf(h)
    ref = h0.copy()
    piv_o, sing_o = oracle.getrf(ref)
    assert sing == sing_o is None
    if piv_s == piv_o:
        assert np.max(np.abs(h - ref)) <= 1e-3 * np.max(np.abs(ref))
    eps = np.finfo(np.float32).eps / 2

    def be(lu, pv):
        perm = devcheck.perm_from_pivots(pv, ms)
        L = np.tril(lu.astype(np.float64), -1)[:, :ns]
        L[np.arange(ns), np.arange(ns)] = 1.0
        U = np.triu(lu.astype(np.float64)[:ns])
        pa = h0.astype(np.float64)[perm]
        return np.linalg.norm(pa - L @ U) / (ms * eps * np.linalg.norm(pa))
    assert be(h, piv_s) <= 10 * max(be(ref, piv_o), 0.01)


def test_c5b_n16384_f64_backward_error_and_lapack_pivots():
    """BASELINE configs[4], f64 n = 16 384: device-side backward error <= 0.5 and the pivot vector of LAPACK's dgetrf
    (same first-maximum rule on continuous data; the oracle itself would need ~25 minutes at this size)."""
    import torch
    import devcheck
    n = 16384
    gen = torch.Generator(device="cuda")
    gen.manual_seed(6)
    a0 = torch.rand(n, n, dtype=torch.float64, device="cuda", generator=gen) * 10
    a = a0.clone()
    ipiv, info = _getrf_dev(a)
    assert info == -1
    piv = ipiv.cpu().numpy()
    assert devcheck.backward_error_dev(a0, a, piv) <= BOUND
    scipy_linalg = pytest.importorskip("scipy.linalg")
    host = a0.cpu().numpy()
    del a0, a
    torch.cuda.empty_cache()
    _, piv_l = scipy_linalg.lu_factor(host, overwrite_a=True, check_finite=False)
    assert np.array_equal(piv, piv_l.astype(piv.dtype))


def test_c3_batched_1e6_f64_bit_exact_sample_and_reconstruction():
    """10^6 x (32x32) f64 on the device: bit-exact against the oracle on a strided sample, P A = L U for every matrix."""
    import torch
    import oracle
    from lair_b200 import _ffi
    batch = 1_000_000
    gen = torch.Generator(device="cuda")
    gen.manual_seed(33)
    a0 = torch.rand(batch, 32, 32, dtype=torch.float64, device="cuda", generator=gen) * 10
    a = a0.clone()
    ipiv = torch.empty(batch, 32, dtype=torch.int32, device="cuda")
    info = torch.empty(batch, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    _ffi.check(_ffi.lib().lair_b200_dgetrf_batched_dev(batch, 32, a.data_ptr(), ipiv.data_ptr(), info.data_ptr(), stream))
    torch.cuda.synchronize()
    assert int((info != -1).sum().item()) == 0
    idx = torch.arange(0, batch, 97, device="cuda")
    ref = a0[idx].cpu().numpy().copy()
    piv_o, _ = oracle.getrf_batched(ref)
    assert np.array_equal(ipiv[idx].cpu().numpy(), piv_o.astype(np.int32))
    assert np.array_equal(a[idx].cpu().numpy(), ref)
    # reconstruction for every matrix, in chunks (f64 on the device)
    worst = 0.0
    for c0 in range(0, batch, 100_000):
        LU = a[c0:c0 + 100_000]
        nb = LU.shape[0]
        L = torch.tril(LU, -1) + torch.eye(32, dtype=torch.float64, device="cuda")
        rec = L @ torch.triu(LU)
        pv = ipiv[c0:c0 + nb].long()
        perm = torch.arange(32, device="cuda").repeat(nb, 1)
        rows = torch.arange(nb, device="cuda")
        for j in range(32):
            p = pv[:, j]
            tmp = perm[rows, j].clone()
            perm[rows, j] = perm[rows, p]
            perm[rows, p] = tmp
        PA = torch.gather(a0[c0:c0 + nb], 1, perm[:, :, None].expand(-1, -1, 32))
        err = torch.linalg.norm((PA - rec).reshape(nb, -1), dim=1) / torch.linalg.norm(PA.reshape(nb, -1), dim=1)
        worst = max(worst, float(err.max()))
    assert worst <= 32 * (np.finfo(np.float64).eps / 2) * 10
