"""GPU: CUDA path vs the CPU oracle on seeded random inputs, through the C ABI.

Bars (north_star): pivot vectors identical on matrices without near-ties; bit-identical
L\\U where the kernel keeps the reference's operation order (batched 32x32, single-CTA small
path on standard layouts); otherwise scaled backward error and solve residual <= 10x the
oracle's own on the same input.
"""
import numpy as np
import pytest

import oracle
from golden_util import backward_error

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lair():
    import lair_b200
    return lair_b200


def _rand(rng, shape, dt, dist="uniform"):
    a = rng.uniform(0, 10, size=shape) if dist == "uniform" else rng.standard_normal(shape)
    if np.issubdtype(dt, np.complexfloating):
        a = a + 1j * (rng.uniform(0, 10, size=shape) if dist == "uniform" else rng.standard_normal(shape))
    return a.astype(dt)


# ---- batched 32x32 (C3): bit-exact ---------------------------------------------------------
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("dist", ["uniform", "normal"])
def test_batched32_bit_exact(lair, dt, dist):
    rng = np.random.default_rng(3)
    a0 = _rand(rng, (4000, 32, 32), dt, dist)
    a = a0.copy()
    ipiv, info = lair.lapack.getrf_batched(a)
    ref = a0.copy()
    piv_o, info_o = oracle.getrf_batched(ref)
    assert np.array_equal(ipiv, piv_o.astype(np.int32))
    assert np.array_equal(info, info_o.astype(np.int32))
    assert np.array_equal(a, ref)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("cfg", [0, 1, 2, 3, 4, 5, 6, 7, 32, 33, 128, 129, 132, 133, 256, 260, 512, 513])
def test_batched32_every_variant_bit_exact(lair, dt, cfg):
    """Every tuning variant of the batched kernel (incl. two matrices per warp, one warp per CTA
    with packed f32x2 updates, retiring rows with NaN-poisoned lanes, odd batch, ties, singular, NaN, infinite and subnormal inputs) is
    bit-identical to the oracle."""
    from lair_b200 import _ffi
    rng = np.random.default_rng(100 + cfg)
    a0 = _rand(rng, (1001, 32, 32), dt)
    a0[5] = rng.integers(-2, 3, size=(32, 32)).astype(dt)
    a0[6] = 0
    a0[7, :, 3] = 0
    a0[8, 4, 4] = np.nan
    a0[9, 2, 2] = np.inf
    a0[10] = a0[10] * dt(1e-39 if dt == np.float32 else 1e-309)  # subnormal pivots: reciprocals overflow
    a0[11] = a0[11] * dt(1e37 if dt == np.float32 else 1e307)    # huge entries: subnormal reciprocals
    a0[1000] = 1
    ref = a0.copy()
    piv_o, info_o = oracle.getrf_batched(ref)
    try:
        _ffi.set_option("batched_cfg", cfg)
        a = a0.copy()
        ipiv, info = lair.lapack.getrf_batched(a)
    finally:
        _ffi.set_option("batched_cfg", -1)
    assert np.array_equal(ipiv, piv_o.astype(np.int32))
    assert np.array_equal(info, info_o.astype(np.int32))
    assert np.array_equal(a, ref, equal_nan=True)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 2, 5, 17, 31])
def test_batched_small_n_bit_exact(lair, dt, n):
    rng = np.random.default_rng(n)
    a0 = _rand(rng, (257, n, n), dt)
    a = a0.copy()
    ipiv, info = lair.lapack.getrf_batched(a)
    ref = a0.copy()
    piv_o, info_o = oracle.getrf_batched(ref)
    assert np.array_equal(ipiv, piv_o.astype(np.int32)) and np.array_equal(info, info_o.astype(np.int32))
    assert np.array_equal(a, ref)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_batched_ties_singular_nan(lair, dt):
    """Exact ties (small integers), singular and rank-deficient matrices, NaN entries, empty batch."""
    rng = np.random.default_rng(11)
    a0 = rng.integers(-2, 3, size=(600, 32, 32)).astype(dt)  # many exact ties and zero pivots
    a0[0] = 0
    a0[1] = 1
    a0[2, :, 5] = 0
    a0[3, 7, 7] = np.nan
    a0[4, :, 0] = np.nan
    a = a0.copy()
    ipiv, info = lair.lapack.getrf_batched(a)
    ref = a0.copy()
    piv_o, info_o = oracle.getrf_batched(ref)
    assert np.array_equal(ipiv, piv_o.astype(np.int32))
    assert np.array_equal(info, info_o.astype(np.int32))
    assert np.array_equal(a, ref, equal_nan=True)
    assert info[0] == 31 and info[1] == 31
    e_ipiv, e_info = lair.lapack.getrf_batched(np.zeros((0, 32, 32), dtype=dt))
    assert e_ipiv.shape == (0, 32) and e_info.shape == (0,)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("chunk", [256, 1000, 8192])
def test_batched_host_entry_chunked_pipeline(lair, dt, chunk):
    """The host-pointer entry pipelines H2D / factor / D2H in chunks of `batched_chunk` matrices and returns pivots and
    info in one piece afterwards: every chunking, ragged last chunk and the > 64-chunk regrouping included, gives the
    oracle's bits."""
    from lair_b200 import _ffi
    rng = np.random.default_rng(23)
    a0 = _rand(rng, (17001, 32, 32), dt)  # chunk 256 -> 67 chunks, regrouped to 64 of 266 with a ragged tail
    a = a0.copy()
    old = _ffi.get_option("batched_chunk")
    _ffi.set_option("batched_chunk", chunk)
    try:
        ipiv, info = lair.lapack.getrf_batched(a)
    finally:
        _ffi.set_option("batched_chunk", old)
    ref = a0.copy()
    piv_o, info_o = oracle.getrf_batched(ref)
    assert np.array_equal(ipiv, piv_o.astype(np.int32))
    assert np.array_equal(info, info_o.astype(np.int32))
    assert np.array_equal(a, ref)


# ---- single-CTA small path (C1 shape): bit-exact on standard layouts ------------------------
@pytest.mark.parametrize("dt", [np.float32, np.float64, np.complex64, np.complex128])
@pytest.mark.parametrize("shape", [(1, 1), (2, 2), (7, 7), (33, 33), (100, 100), (128, 128), (40, 17), (17, 40), (128, 3)])
def test_small_bit_exact_row_major(lair, dt, shape):
    rng = np.random.default_rng(shape[0] * 131 + shape[1])
    a0 = _rand(rng, shape, dt)
    a = a0.copy()
    piv, sing = lair.lapack.getrf(a)
    ref = a0.copy()
    piv_o, sing_o = oracle.getrf(ref)
    assert piv == piv_o and sing == sing_o
    assert np.array_equal(a, ref)


@pytest.mark.parametrize("dt", [np.float64, np.complex128, np.float32])
def test_small_other_layouts_match_reference(lair, dt):
    """Non-standard layouts take the reference's left-looking body: same pivots, LU within rounding."""
    rng = np.random.default_rng(5)
    a0 = _rand(rng, (100, 100), dt)  # benches/getrf.rs shape
    for make in (np.asfortranarray, lambda x: np.ascontiguousarray(x[::-1, ::-1])[::-1, ::-1],
                 lambda x: np.repeat(np.repeat(x, 2, axis=0), 3, axis=1)[::2, ::3]):
        a = make(a0.copy())
        ref = make(a0.copy())
        assert np.array_equal(a, a0)
        piv, sing = lair.lapack.getrf(a)
        piv_o, sing_o = oracle.getrf(ref)
        assert piv == piv_o and sing == sing_o
        tol = 1e3 * np.finfo(dt).eps * 10 * 100
        assert np.max(np.abs(a - ref)) <= tol
        assert backward_error(a0, a, piv) <= 10 * max(backward_error(a0, ref, piv_o), 0.01)


def test_small_singular_semantics(lair):
    for a0 in (np.zeros((3, 3)), np.array([[1.0, 2, 3], [2, 4, 6], [3, 6, 9]]), np.ones((2, 2))):
        a, ref = a0.copy(), a0.copy()
        piv, sing = lair.lapack.getrf(a)
        piv_o, sing_o = oracle.getrf(ref)
        assert (piv, sing) == (piv_o, sing_o)
        assert np.array_equal(a, ref)
    e = np.zeros((0, 0), dtype=np.float32)
    assert lair.lapack.getrf(e) == ([], None)
    assert lair.lapack.getrf(np.zeros((0, 5))) == ([], None)


def test_small_getrs_bit_exact(lair):
    rng = np.random.default_rng(9)
    for n in (1, 2, 17, 100, 128):
        for dt in (np.float32, np.float64, np.complex128):
            a = _rand(rng, (n, n), dt)
            b = _rand(rng, (n,), dt)
            piv, _ = oracle.getrf(a)
            x = lair.lapack.getrs(a, piv, b)
            assert np.array_equal(x, oracle.getrs(a, piv, b)), (n, dt)
            # strided b, transposed-layout factors
            x2 = lair.lapack.getrs(np.asfortranarray(a), piv, np.repeat(b, 2)[::2])
            assert np.array_equal(x, x2)


# ---- building blocks of the blocked path ---------------------------------------------------
def _dev(arr):
    import torch
    return torch.from_numpy(np.ascontiguousarray(arr)).cuda()


def _stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("dt,pfx", [(np.float64, "d"), (np.float32, "s")])
def test_laswp_dev_matches_oracle(lair, dt, pfx):
    import torch
    from lair_b200 import _ffi
    rng = np.random.default_rng(21)
    for rows, ncols, k0, k1 in ((300, 70, 0, 32), (1000, 513, 32, 300), (64, 1, 0, 64), (5000, 129, 100, 356), (777, 35, 10, 106)):  # (last: 96 interchanges on an unaligned f64 tail = 48 KB of staging + the static tables)
        a0 = rng.standard_normal((rows, ncols)).astype(dt)
        piv = np.arange(rows, dtype=np.int64)
        for i in range(k0, k1):
            piv[i] = rng.integers(i, rows)
        piv[k0 + 1] = piv[k0]  # a repeated target
        ref = a0.copy()
        oracle.laswp(ref, piv[:k1], begin=k0)
        d = _dev(a0)
        dp = torch.from_numpy(piv.astype(np.int32)).cuda()
        fn = getattr(_ffi.lib(), f"lair_b200_{pfx}laswp_dev")
        _ffi.check(fn(ncols, d.data_ptr(), ncols, k0, k1, dp.data_ptr(), _stream()))
        torch.cuda.synchronize()
        assert np.array_equal(d.cpu().numpy(), ref), (rows, ncols, k0, k1)


@pytest.mark.parametrize("dt,pfx", [(np.float64, "d"), (np.float32, "s")])
def test_gemm_minus_dev(lair, dt, pfx):
    import torch
    from lair_b200 import _ffi
    rng = np.random.default_rng(22)
    for m, n, k, pad in ((128, 128, 16, 0), (257, 130, 48, 0), (1000, 777, 256, 0), (64, 40, 33, 1), (300, 200, 32, 3)):
        lda, ldb, ldc = k + pad, n + pad, n + pad
        a = rng.standard_normal((m, lda)).astype(dt)
        b = rng.standard_normal((k, ldb)).astype(dt)
        c = rng.standard_normal((m, ldc)).astype(dt)
        exp = c.astype(np.float64).copy()
        exp[:, :n] -= a[:, :k].astype(np.float64) @ b[:, :n].astype(np.float64)
        da, db, dc = _dev(a), _dev(b), _dev(c)
        fn = getattr(_ffi.lib(), f"lair_b200_{pfx}gemm_minus_dev")
        _ffi.check(fn(m, n, k, da.data_ptr(), lda, db.data_ptr(), ldb, dc.data_ptr(), ldc, _stream()))
        torch.cuda.synchronize()
        got = dc.cpu().numpy()
        tol = 50 * np.finfo(dt).eps * np.sqrt(k) * 4
        assert np.max(np.abs(got[:, :n] - exp[:, :n])) <= tol * max(1.0, np.max(np.abs(exp))), (m, n, k)
        assert np.array_equal(got[:, n:], c[:, n:])  # padding untouched


@pytest.mark.parametrize("dt,pfx", [(np.float64, "d"), (np.float32, "s")])
def test_trsm_dev_matches_oracle(lair, dt, pfx):
    import torch
    from lair_b200 import _ffi
    rng = np.random.default_rng(23)
    for k, ncols in ((32, 100), (20, 7), (256, 300), (96, 1000)):
        l = np.tril(rng.uniform(-1, 1, (k, k)), -1).astype(dt) / max(1, k // 8) + np.eye(k, dtype=dt) * 3  # diagonal ignored (unit)
        b = rng.standard_normal((k, ncols)).astype(dt)
        ref = b.astype(np.float64).copy()
        lib = oracle.lib()
        l64 = l.astype(np.float64)
        lib.oracle_dtrsm(l64.ctypes.data, k, 1, ref.ctypes.data, k, ncols, ncols, 1)
        dl, db = _dev(l), _dev(b)
        fn = getattr(_ffi.lib(), f"lair_b200_{pfx}trsm_dev")
        _ffi.check(fn(k, ncols, dl.data_ptr(), k, db.data_ptr(), ncols, _stream()))
        torch.cuda.synchronize()
        got = db.cpu().numpy().astype(np.float64)
        assert np.max(np.abs(got - ref)) <= 1e4 * np.finfo(dt).eps * max(1.0, np.max(np.abs(ref))), (k, ncols)


# ---- blocked factorization vs the oracle -----------------------------------------------------
def _first_divergence(p, q):
    for i, (x, y) in enumerate(zip(p, q)):
        if x != y:
            return i
    return None


@pytest.mark.parametrize("shape", [(129, 129), (200, 200), (500, 500), (1000, 1000), (1024, 1024), (2048, 2048),
                                   (3000, 200), (200, 700), (777, 333), (333, 777)])
def test_blocked_f64_matches_oracle(lair, shape):
    rng = np.random.default_rng(shape[0] + 7 * shape[1])
    a0 = _rand(rng, shape, np.float64)
    a = a0.copy()
    piv, sing = lair.lapack.getrf(a)
    ref = a0.copy()
    piv_o, sing_o = oracle.getrf(ref)
    assert sing == sing_o is None
    assert piv == piv_o, f"first divergence at step {_first_divergence(piv, piv_o)}"
    scale = np.max(np.abs(ref))
    assert np.max(np.abs(a - ref)) <= 1e-9 * scale
    be, be_o = backward_error(a0, a, piv), backward_error(a0, ref, piv_o)
    assert be <= 10 * max(be_o, 0.01), (be, be_o)


@pytest.mark.parametrize("option,values", [("panel_cluster", (0, 1, 2, 3)), ("panel_rpt", (1, 4, 2, 0)), ("lookahead", (0, 1)), ("gemm_cfg", (1, 2, 3, 0)),
                                           ("nb", (64, 128, 512, 256)), ("fuse_swap_trsm", (0, 2, 1)), ("panel_exchange", (0, 1)), ("panel_w64", (0, 1)), ("chain_on_p", (0, 1)), ("trsm_strip", (0, 1, 2))])
def test_blocked_f64_kernel_variants(lair, option, values):
    """Every kernel variant behind a tuning option produces the oracle's pivots and L\\U."""
    from lair_b200 import _ffi
    rng = np.random.default_rng(4242)
    a0 = _rand(rng, (1100, 900), np.float64)
    ref = a0.copy()
    piv_o, sing_o = oracle.getrf(ref)
    default = _ffi.get_option(option)
    try:
        for v in values:
            _ffi.set_option(option, v)
            a = a0.copy()
            piv, sing = lair.lapack.getrf(a)
            assert piv == piv_o and sing == sing_o, (option, v, _first_divergence(piv, piv_o))
            assert np.max(np.abs(a - ref)) <= 1e-9 * np.max(np.abs(ref)), (option, v)
    finally:
        _ffi.set_option(option, default)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_panel_generations_bit_identical(lair, dt):
    """The cluster panel kernels behind `panel_cluster` (3: the pushed record carries the candidate row's register window,
    panel_push.cu; 2: winner row pulled over DSMEM, panel_blocked.cu) keep the same operation order: pivots, singular step and
    L\\U agree bit for bit on single-panel shapes -- ragged widths, one CTA up to 16, integer ties, a zero column, a dependent
    column, NaN / inf and subnormal entries -- for every rows-per-thread variant, and the pivots are the oracle's."""
    from lair_b200 import _ffi
    rng = np.random.default_rng(77)
    rows = [200, 513, 4096, 8192] + ([12000] if dt == np.float32 else [])
    d_cluster, d_rpt = _ffi.get_option("panel_cluster"), _ffi.get_option("panel_rpt")
    try:
        for m in rows:
            for w in (7, 32, 33, 64):
                for kind in ("rand", "ties", "zerocol", "nonfinite", "subnormal"):
                    if kind != "rand" and m not in (513, 8192):
                        continue
                    a0 = _rand(rng, (m, w), dt)
                    if kind == "ties":
                        a0 = rng.integers(-3, 4, size=(m, w)).astype(dt)
                    elif kind == "zerocol":
                        a0[:, w // 2] = 0
                        a0[:, 1] = a0[:, 0] * 2
                    elif kind == "nonfinite":
                        a0[m // 3, w // 3] = np.nan
                        a0[m - 1, 0] = np.inf
                    elif kind == "subnormal":
                        a0[:, 2] *= dt(1e-42 if dt == np.float32 else 1e-312)
                    _ffi.set_option("panel_cluster", 2)
                    _ffi.set_option("panel_rpt", 2)
                    ref = a0.copy()
                    with np.errstate(all="ignore"):
                        piv_r, sing_r = lair.lapack.getrf(ref)
                    if kind == "rand" and m <= 4096:  # (the other kinds are near-tie factories: FMA vs mul + sub may pick differently)
                        orc = a0.copy()
                        piv_o, sing_o = oracle.getrf(orc)
                        assert piv_r == piv_o and sing_r == sing_o, (m, w, kind)
                    _ffi.set_option("panel_cluster", 3)
                    for rpt in (0, 1, 2, 4):
                        _ffi.set_option("panel_rpt", rpt)
                        a = a0.copy()
                        piv, sing = lair.lapack.getrf(a)
                        assert piv == piv_r and sing == sing_r, (m, w, kind, rpt, _first_divergence(piv, piv_r))
                        assert a.tobytes() == ref.tobytes(), (m, w, kind, rpt)
    finally:
        _ffi.set_option("panel_cluster", d_cluster)
        _ffi.set_option("panel_rpt", d_rpt)


@pytest.mark.parametrize("shape", [(1300, 1300), (1500, 1100), (1100, 1700)])
def test_blocked_f64_chunked_upload(lair, shape):
    """Host-pointer getrf with the matrix uploaded in column chunks while the sweep already runs
    (late chunks catch up with one laswp + trsm + gemm): oracle pivots and L\\U, and the same
    solution through gesv."""
    from lair_b200 import _ffi
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    a0 = _rand(rng, shape, np.float64)
    ref = a0.copy()
    piv_o, sing_o = oracle.getrf(ref)
    default = _ffi.get_option("stream_cols")
    try:
        for w in (512, 0):
            _ffi.set_option("stream_cols", w)
            a = a0.copy()
            piv, sing = lair.lapack.getrf(a)
            assert piv == piv_o and sing == sing_o, (w, _first_divergence(piv, piv_o))
            assert np.max(np.abs(a - ref)) <= 1e-9 * np.max(np.abs(ref)), w
            if shape[0] == shape[1]:
                b = _rand(rng, (shape[0], 3), np.float64)
                x = lair.equation.solve(a0, b)
                n = shape[0]
                res = np.linalg.norm(a0 @ x - b) / (np.linalg.norm(a0) * np.linalg.norm(x) * n * np.finfo(np.float64).eps)
                assert res < 1.0, (w, res)
    finally:
        _ffi.set_option("stream_cols", default)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_trsm_strip_bit_identical_to_fused_chain(lair, dt):
    """The one-launch register-tiled triangle solve of the wide trailing ranges (trsm_strip.cu, after one laswp pass) applies
    the same FMAs in the same order as the chain of fused 64-row launches (laswp_trsm.cu): pivots and L\\U byte-identical
    for every block width, including widths that leave ragged L blocks (96, 160, 224) and a ragged last block."""
    from lair_b200 import _ffi
    rng = np.random.default_rng(5150)
    d_strip, d_nb = _ffi.get_option("trsm_strip"), _ffi.get_option("nb")
    try:
        for shape in ((1777, 1777), (1500, 2300)):
            a0 = _rand(rng, shape, dt)
            for nb in (0, 96, 160, 224, 256):
                _ffi.set_option("nb", nb)
                _ffi.set_option("trsm_strip", 0)
                ref = a0.copy()
                piv_r, sing_r = lair.lapack.getrf(ref)
                for strip in (1, 2):
                    _ffi.set_option("trsm_strip", strip)
                    a = a0.copy()
                    piv, sing = lair.lapack.getrf(a)
                    assert piv == piv_r and sing == sing_r, (shape, nb, strip, _first_divergence(piv, piv_r))
                    assert a.tobytes() == ref.tobytes(), (shape, nb, strip)
    finally:
        _ffi.set_option("trsm_strip", d_strip)
        _ffi.set_option("nb", d_nb)


@pytest.mark.parametrize("shape", [(4096, 4096), (5000, 5000), (5300, 4200), (3900, 5100)])
def test_paired_k512_update_bit_identical(lair, shape):
    """While many columns remain, two 256-wide block steps share one K = 512 trailing GEMM (blocked.cu pair_update): the
    second block's rows take the first block's contribution and their own solve, the rows below one GEMM over both column
    blocks.  Same FMAs in the same order: pivots and L\\U byte-identical to the step-by-step sweep (thresholds lowered so
    that the pairing is active at test sizes; with and without the lookahead chain on the panel stream)."""
    from lair_b200 import _ffi
    rng = np.random.default_rng(shape[0] + shape[1])
    a0 = _rand(rng, shape, np.float64)
    saved = {k: _ffi.get_option(k) for k in ("pair_k512", "nb_t1", "nb_t2", "chain_on_p")}
    try:
        _ffi.set_option("nb_t2", 1024)   # 256-wide blocks while more than 1024 columns remain
        _ffi.set_option("nb_t1", 512)
        _ffi.set_option("pair_k512", 0)
        ref = a0.copy()
        piv_r, sing_r = lair.lapack.getrf(ref)
        for cop in (saved["chain_on_p"], 0, 1 << 30):
            _ffi.set_option("chain_on_p", cop)
            _ffi.set_option("pair_k512", 1200)
            a = a0.copy()
            piv, sing = lair.lapack.getrf(a)
            assert piv == piv_r and sing == sing_r, (cop, _first_divergence(piv, piv_r))
            assert a.tobytes() == ref.tobytes(), cop
    finally:
        for k, v in saved.items():
            _ffi.set_option(k, v)


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("shape", [(3000, 3000), (2500, 4100), (4100, 2300), (6144, 6144)])
def test_getrf_pinned_rows_drain_while_sweeping(lair, dt, shape):
    """Host-pointer getrf on a PINNED row-contiguous array sends finished rows home while the sweep still runs
    (RowDrain, blocked.cu): the array and the pivots must be byte-identical to the copy-at-the-end path, with and
    without the chunked upload, also through a padded row stride."""
    import torch
    from lair_b200 import _ffi
    rng = np.random.default_rng(shape[0] * 3 + shape[1])
    a0 = _rand(rng, shape, dt)
    m, n = shape
    d_drain, d_cols = _ffi.get_option("drain_rows"), _ffi.get_option("stream_cols")
    try:
        for cols in (d_cols, 0):
            _ffi.set_option("stream_cols", cols)
            _ffi.set_option("drain_rows", 0)  # (the reference per upload mode: late column chunks catch up through other kernels)
            ref = a0.copy()
            piv_r, sing_r = lair.lapack.getrf(ref)
            for pad in (0, 24):
                _ffi.set_option("drain_rows", 1)
                buf = torch.empty((m, n + pad), dtype=torch.float64 if dt == np.float64 else torch.float32, pin_memory=True).numpy()
                buf[:] = -7
                a = buf[:, :n]
                a[:] = a0
                piv, sing = lair.lapack.getrf(a)
                assert piv == piv_r and sing == sing_r, (cols, pad, _first_divergence(piv, piv_r))
                assert np.array_equal(a, ref), (cols, pad)
                if pad:
                    assert np.all(buf[:, n:] == -7)  # the padding columns of the host array are never written
    finally:
        _ffi.set_option("drain_rows", d_drain)
        _ffi.set_option("stream_cols", d_cols)


def test_getrf_pinned_feed_drain_and_pairing_together(lair):
    """The host-pointer getrf with everything on at once -- column chunks arriving during the sweep, finished rows draining
    to the pinned array, two block steps sharing one trailing GEMM (thresholds lowered to test sizes) -- returns the bytes of
    the plain path (same upload mode, no drain, no pairing)."""
    import torch
    from lair_b200 import _ffi
    rng = np.random.default_rng(808)
    m = n = 5200
    a0 = _rand(rng, (m, n), np.float64)
    saved = {k: _ffi.get_option(k) for k in ("pair_k512", "nb_t1", "nb_t2", "drain_rows", "stream_cols")}
    try:
        _ffi.set_option("nb_t2", 1024)
        _ffi.set_option("nb_t1", 512)
        for cols in (saved["stream_cols"], 0):
            _ffi.set_option("stream_cols", cols)
            _ffi.set_option("pair_k512", 0)
            _ffi.set_option("drain_rows", 0)
            ref = torch.empty((m, n), dtype=torch.float64, pin_memory=True).numpy()
            ref[:] = a0
            piv_r, sing_r = lair.lapack.getrf(ref)
            _ffi.set_option("pair_k512", 1200)
            _ffi.set_option("drain_rows", 1)
            a = torch.empty((m, n), dtype=torch.float64, pin_memory=True).numpy()
            a[:] = a0
            piv, sing = lair.lapack.getrf(a)
            assert piv == piv_r and sing == sing_r, (cols, _first_divergence(piv, piv_r))
            assert a.tobytes() == ref.tobytes(), cols
    finally:
        for k, v in saved.items():
            _ffi.set_option(k, v)


@pytest.mark.parametrize("shape", [(1300, 1300), (1100, 1700)])
def test_blocked_f32_chunked_upload_backward_error(lair, shape):
    """f32 through the chunked-upload host path (late chunks catch up with laswp + recursive trsm +
    FFMA gemm): backward error within 10x the oracle's."""
    from lair_b200 import _ffi
    rng = np.random.default_rng(shape[0] + shape[1])
    a0 = _rand(rng, shape, np.float32, "normal")
    ref = a0.copy()
    piv_o, _ = oracle.getrf(ref)
    be_o = backward_error(a0, ref, piv_o)
    default = _ffi.get_option("stream_cols")
    try:
        for w in (512, 0):
            _ffi.set_option("stream_cols", w)
            a = a0.copy()
            piv, sing = lair.lapack.getrf(a)
            assert sing is None
            be = backward_error(a0, a, piv)
            assert be <= 10 * max(be_o, 0.01), (w, be, be_o)
    finally:
        _ffi.set_option("stream_cols", default)


def _oracle_column_at_step(a0, lu_o, piv_o, d):
    """Rows d.. of the column the oracle searched at step d, rebuilt in f64 from its own factors:
    (P_d A)[d:, d] - L_d[d:, :d] U[:d, d], where P_d holds the first d interchanges and L_d is the oracle's L with the
    interchanges of the steps >= d undone (getrf.rs:65-70 swaps whole rows, so multipliers travel with their rows)."""
    m = a0.shape[0]
    perm = np.arange(m)
    for i in range(d):
        p = piv_o[i]
        perm[i], perm[p] = perm[p], perm[i]
    lu64 = lu_o.astype(np.float64)
    Ld = np.tril(lu64, -1)[:, :d].copy()
    for i in range(len(piv_o) - 1, d - 1, -1):
        p = piv_o[i]
        if p != i:
            Ld[[i, p]] = Ld[[p, i]]
    return a0.astype(np.float64)[perm, d][d:] - Ld[d:] @ lu64[:d, d]


@pytest.mark.parametrize("shape", [(300, 300), (1000, 1000), (2000, 300), (300, 900)])
def test_blocked_f32_matches_oracle(lair, shape):
    rng = np.random.default_rng(shape[0] + 13 * shape[1])
    a0 = _rand(rng, shape, np.float32)
    a = a0.copy()
    piv, sing = lair.lapack.getrf(a)
    ref = a0.copy()
    piv_o, sing_o = oracle.getrf(ref)
    assert sing == sing_o is None
    d = _first_divergence(piv, piv_o)
    if d is not None:
        # The pivot vectors may differ only at a NEAR-TIE (north star: "identical on matrices without near-ties"):
        # rebuild the column the oracle searched at step d from its own factors -- rows d.. of P_d A - L[:, :d] U[:d, d],
        # in f64 -- and require our candidate to be within accumulated f32 rounding of the oracle's maximum.
        s = _oracle_column_at_step(a0, ref, piv_o, d)
        ours, best = abs(s[piv[d] - d]), np.max(np.abs(s))
        assert abs(s[piv_o[d] - d]) >= best * (1 - 5e-4), "the oracle's own pivot is not (nearly) the maximum of the rebuilt column"
        assert ours >= best * (1 - 5e-4), f"pivot divergence at step {d} is not a near-tie: |ours| = {ours}, max = {best}"
    be, be_o = backward_error(a0, a, piv), backward_error(a0, ref, piv_o)
    assert be <= 10 * max(be_o, 0.01), (be, be_o)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(300, 300), (1500, 1500), (2500, 400)])
def test_blocked_backward_error(lair, dt, shape):
    rng = np.random.default_rng(99)
    a0 = _rand(rng, shape, dt, "normal")
    a = a0.copy()
    piv, sing = lair.lapack.getrf(a)
    ref = a0.copy()
    piv_o, _ = oracle.getrf(ref)
    assert sing is None
    assert sorted(set(piv)) is not None
    be, be_o = backward_error(a0, a, piv), backward_error(a0, ref, piv_o)
    assert be <= 10 * max(be_o, 0.01), (be, be_o)


def test_blocked_layouts_and_singular(lair):
    rng = np.random.default_rng(17)
    a0 = _rand(rng, (400, 400), np.float64)
    ref = a0.copy()
    piv_o, _ = oracle.getrf(ref)
    for make in (np.asfortranarray, lambda x: np.ascontiguousarray(x[::-1, ::-1])[::-1, ::-1]):
        a = make(a0.copy())
        piv, sing = lair.lapack.getrf(a)
        assert piv == piv_o and sing is None
        assert np.max(np.abs(a - ref)) <= 1e-9 * np.max(np.abs(ref))
    # rank-deficient: two equal columns -> a zero pivot late in the factorization
    s0 = a0.copy()
    s0[:, 300] = s0[:, 10]
    s, sref = s0.copy(), s0.copy()
    piv, sing = lair.lapack.getrf(s)
    piv_o2, sing_o2 = oracle.getrf(sref)
    # exact cancellation is rounding dependent; both must agree that the factorization completes
    assert len(piv) == len(piv_o2)
    z = np.zeros((300, 300))
    z[:150, :150] = a0[:150, :150]
    zz, zref = z.copy(), z.copy()
    piv, sing = lair.lapack.getrf(zz)
    piv_o3, sing_o3 = oracle.getrf(zref)
    assert sing == sing_o3 == 299
    assert piv == piv_o3


@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("n,nrhs", [(300, 1), (300, 5), (1000, 64), (2048, 64)])
def test_blocked_getrs_matches_oracle(lair, dt, n, nrhs):
    rng = np.random.default_rng(n + nrhs)
    a0 = _rand(rng, (n, n), dt)
    b = _rand(rng, (n, nrhs), dt)
    lu = a0.copy()
    piv, _ = oracle.getrf(lu)
    x = lair.lapack.getrs(lu, piv, b)
    assert x.shape == (n, nrhs)
    eps = np.finfo(dt).eps / 2
    a64, b64 = a0.astype(np.float64), b.astype(np.float64)
    for r in range(0, nrhs, max(1, nrhs // 4)):
        xo = oracle.getrs(lu, piv, np.ascontiguousarray(b[:, r]))
        res = np.linalg.norm(a64 @ x[:, r].astype(np.float64) - b64[:, r]) / (np.linalg.norm(a64) * np.linalg.norm(x[:, r]) * n * eps)
        res_o = np.linalg.norm(a64 @ xo.astype(np.float64) - b64[:, r]) / (np.linalg.norm(a64) * np.linalg.norm(xo) * n * eps)
        assert res <= 10 * max(res_o, 0.01), (res, res_o)
    one = lair.lapack.getrs(lu, piv, np.ascontiguousarray(b[:, 0]))
    assert one.shape == (n,) and np.allclose(one, x[:, 0], rtol=0, atol=1e3 * np.finfo(dt).eps * np.max(np.abs(x)))


@pytest.mark.parametrize("n,nrhs", [(200, 3), (777, 5), (1000, 64), (1030, 65), (2048, 130)])
def test_getrs_dataflow_vs_recursive(lair, n, nrhs):
    """The persistent dataflow triangular solves (flag-in-data and flag-word variants) and the
    recursive TRSM agree with the oracle."""
    from lair_b200 import _ffi
    rng = np.random.default_rng(n * 3 + nrhs)
    a0 = _rand(rng, (n, n), np.float64)
    b = _rand(rng, (n, nrhs), np.float64)
    lu = a0.copy()
    piv, _ = oracle.getrf(lu)
    xo = np.stack([oracle.getrs(lu, piv, np.ascontiguousarray(b[:, r])) for r in range(0, nrhs, max(1, nrhs // 3))], axis=1)
    try:
        for df, rb, perm in ((2, 32, 1), (2, 32, 0), (1, 64, 1), (1, 32, 0), (0, 32, 1)):
            _ffi.set_option("trsm_dataflow", df)
            _ffi.set_option("trsm_rb", rb)
            _ffi.set_option("laswp_perm", perm)  # P b as one collapsed permutation vs pass-by-pass interchanges
            x = lair.lapack.getrs(lu, piv, b)
            got = x[:, ::max(1, nrhs // 3)]
            assert np.max(np.abs(got - xo)) <= 1e-7 * np.max(np.abs(xo)), (df, rb, perm, np.max(np.abs(got - xo)))
    finally:
        _ffi.set_option("trsm_dataflow", 2)
        _ffi.set_option("trsm_rb", 32)
        _ffi.set_option("laswp_perm", 1)


def test_no_device_fault_after_tall_panel_and_solves(lair):
    """The bounded cross-CTA waits (tall-panel exchange, dataflow solves) did not time out: the fault
    word the host entry points check before returning is clear after device-resident calls too."""
    import torch
    from lair_b200 import _ffi
    L = _ffi.lib()
    m, n = 20000, 64  # taller than one cluster: the global-memory exchange panel
    a = torch.rand(m, n, dtype=torch.float64, device="cuda")
    ipiv = torch.empty(n, dtype=torch.int32, device="cuda")
    info = torch.empty(1, dtype=torch.int32, device="cuda")
    _ffi.check(L.lair_b200_dgetrf_dev(m, n, a.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), _stream()))
    n2 = 1500
    lu = torch.rand(n2, n2, dtype=torch.float64, device="cuda")
    piv2 = torch.empty(n2, dtype=torch.int32, device="cuda")
    _ffi.check(L.lair_b200_dgetrf_dev(n2, n2, lu.data_ptr(), n2, piv2.data_ptr(), info.data_ptr(), _stream()))
    b = torch.rand(n2, 70, dtype=torch.float64, device="cuda")
    _ffi.check(L.lair_b200_dgetrs_dev(n2, 70, lu.data_ptr(), n2, piv2.data_ptr(), b.data_ptr(), 70, _stream()))
    _ffi.check_fault(_stream())
    assert int(info.item()) == -1


def test_device_fault_is_reported_loudly(lair):
    """A raised device fault (what a timed-out cross-CTA wait leaves behind) turns the next check --
    explicit, or the one every host-pointer entry point makes before returning -- into an error,
    and is cleared by it."""
    from lair_b200 import _ffi
    _ffi.check_fault()
    _ffi.set_option("debug_raise_fault", 2)
    with pytest.raises(_ffi.LairB200Error, match="timed out"):
        _ffi.check_fault()
    _ffi.check_fault()  # reported once, then clear
    rng = np.random.default_rng(77)
    a0 = _rand(rng, (300, 300), np.float64)
    _ffi.set_option("debug_raise_fault", 1)
    with pytest.raises(_ffi.LairB200Error, match="timed out"):
        lair.lapack.getrf(a0.copy())
    a = a0.copy()
    piv, sing = lair.lapack.getrf(a)  # the library keeps working afterwards
    ref = a0.copy()
    piv_o, sing_o = oracle.getrf(ref)
    assert piv == piv_o and sing == sing_o


def test_equation_solve_end_to_end(lair):
    rng = np.random.default_rng(2)
    n = 1500
    a = _rand(rng, (n, n), np.float64)
    b = _rand(rng, (n,), np.float64)
    a_keep = a.copy()
    x = lair.equation.solve(a, b)
    assert np.array_equal(a, a_keep)  # equation::solve factors a copy (equation.rs:54)
    res = np.linalg.norm(a @ x - b) / (np.linalg.norm(a) * np.linalg.norm(x) * n * np.finfo(np.float64).eps)
    assert res < 1.0


# ---- device-resident lu::Factorized: handle API ------------------------
def _host_plu(lu, piv):
    m, n = lu.shape
    k = min(m, n)
    perm = np.arange(m)
    for i, p in enumerate(piv):  # laswp on the identity (lu.rs:28-39)
        perm[i], perm[p] = perm[p], perm[i]
    P = np.zeros((m, m), dtype=lu.dtype)
    P[perm, np.arange(m)] = 1
    L = np.tril(lu[:, :k], -1)
    L[np.arange(k), np.arange(k)] = 1
    U = np.triu(lu[:k, :])
    PL = np.zeros_like(L)
    PL[perm] = L
    return P, L, U, PL


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.complex64, np.complex128])
@pytest.mark.parametrize("shape", [(1, 1), (5, 5), (100, 100), (40, 17), (17, 40), (128, 3)])
def test_lu_handle_small_matches_getrf(lair, dt, shape):
    """Factorized keeps the factors on the device: same pivots / L\\U / singular flag as getrf on the same
    input (bit for bit), views equal to the host restatement of lu.rs:28-72,107-153, input untouched."""
    rng = np.random.default_rng(shape[0] * 17 + shape[1])
    a0 = _rand(rng, shape, dt)
    a_in = a0.copy()
    f = lair.decomposition.lu.Factorized.from_(a_in)
    assert np.array_equal(a_in, a0)
    ref = a0.copy()
    piv, sing = lair.lapack.getrf(ref)
    assert f.pivots == piv and f.singular == sing and f.is_singular() == (sing is not None)
    assert np.array_equal(f.lu, ref)
    P, L, U, PL = _host_plu(ref, piv)
    assert np.array_equal(f.p(), P) and np.array_equal(f.l(), L) and np.array_equal(f.u(), U)
    k = min(shape)
    assert np.array_equal(f.into_pl()[:, :k], PL)
    if shape[0] == shape[1]:
        b = _rand(rng, (shape[0],), dt)
        assert np.array_equal(f.solve(b), lair.lapack.getrs(ref, piv, b))
        bm = _rand(rng, (shape[0], 3), dt)
        assert np.array_equal(f.solve(bm), lair.lapack.getrs(ref, piv, bm))


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(700, 700), (900, 400), (400, 900)])
def test_lu_handle_blocked_path(lair, dt, shape):
    rng = np.random.default_rng(shape[0] + shape[1])
    a0 = _rand(rng, shape, dt)
    f = lair.decomposition.lu.Factorized.from_(a0 if shape[0] != 900 else np.asfortranarray(a0))
    ref = a0.copy()
    piv, sing = lair.lapack.getrf(ref)
    assert f.pivots == piv and f.singular == sing
    if shape[0] != 900:
        assert np.array_equal(f.lu, ref)  # same kernels, same layout path
    lu = f.lu
    P, L, U, PL = _host_plu(lu, piv)
    assert np.array_equal(f.p(), P) and np.array_equal(f.l(), L) and np.array_equal(f.u(), U)
    assert np.array_equal(f.into_pl()[:, :min(shape)], PL)
    # P^T A = L U within the backward-error bar of the blocked path
    assert backward_error(a0, lu, piv) <= 10 * max(backward_error(a0, ref, piv), 0.01)
    if shape[0] == shape[1]:
        bm = _rand(rng, (shape[0], 5), dt)
        x = f.solve(bm)
        assert np.array_equal(x, lair.lapack.getrs(ref, piv, bm))
        x1 = f.solve(np.ascontiguousarray(bm[:, 2]))
        assert np.max(np.abs(x1 - x[:, 2])) <= 1e-5 * np.max(np.abs(x[:, 2]))  # 1-RHS and multi-RHS solvers differ in order


def test_lu_handle_errors_and_edges(lair):
    from lair_b200 import _ffi
    from lair_b200.errors import InvalidInput
    F = lair.decomposition.lu.Factorized
    f = F.from_(np.zeros((0, 0)))
    assert f.pivots == [] and not f.is_singular() and f.lu.shape == (0, 0) and f.p().shape == (0, 0)
    f = F.from_(np.array([[1.0, 1.0], [1.0, 1.0]]))  # getrf.rs:343-347
    assert f.singular == 1 and f.is_singular()
    f = F.from_(np.arange(12, dtype=np.float64).reshape(4, 3) + np.eye(4, 3))
    with pytest.raises(InvalidInput.Shape):
        f.solve(np.zeros(3))
    with pytest.raises(_ffi.LairB200Error):
        f.solve(np.zeros(4))  # not square: getrs.rs:18-20
    g = F.from_(np.eye(3))
    assert np.array_equal(g.solve(np.array([1.0, 2.0, 3.0])), np.array([1.0, 2.0, 3.0]))
    # a WIDE factorization solves with its leading m x m block (getrs.rs:18-20: a.ncols() >= p.len())
    rng = np.random.default_rng(5)
    w0 = rng.uniform(0, 10, size=(3, 5))
    bw = rng.uniform(0, 10, size=3)
    fw = F.from_(w0.copy())
    ref = w0.copy()
    piv_o, _ = oracle.getrf(ref)
    assert np.array_equal(fw.solve(bw), oracle.getrs(ref, piv_o, bw))
    # pivots that are not sequential interchanges (entry above its step) are refused, not mis-applied
    with pytest.raises(_ffi.LairB200Error):
        lair.lapack.getrs(np.eye(3), [2, 0, 2], np.ones(3))
    # several factorizations coexist
    hs = [F.from_(np.eye(4) * (i + 1)) for i in range(5)]
    for i, h in enumerate(hs):
        assert np.allclose(h.solve(np.ones(4)), 1.0 / (i + 1))


# ---- complex beyond the single-CTA limit: blocked sweep, ZGEMM as one real GEMM (blocked_cx.cu) ----
@pytest.mark.parametrize("dt", [np.complex128, np.complex64])
@pytest.mark.parametrize("shape", [(129, 129), (200, 200), (500, 500), (1000, 1000), (1500, 200), (200, 700), (777, 333), (333, 777)])
def test_complex_blocked_matches_oracle(lair, dt, shape):
    rng = np.random.default_rng(shape[0] * 3 + shape[1])
    a0 = _rand(rng, shape, dt)
    a = a0.copy()
    piv, sing = lair.lapack.getrf(a)
    ref = a0.copy()
    piv_o, sing_o = oracle.getrf(ref)
    assert sing == sing_o is None
    be, be_o = backward_error(a0, a, piv), backward_error(a0, ref, piv_o)
    assert be <= 10 * max(be_o, 0.01), (be, be_o)
    d = _first_divergence(piv, piv_o)
    if dt == np.complex128:
        assert d is None, f"first divergence at step {d}"
        assert np.max(np.abs(a - ref)) <= 1e-9 * np.max(np.abs(ref))
    elif d is None:  # Complex<f32>: identical pivots unless a near-tie (then the backward error above carries parity)
        assert np.max(np.abs(a - ref)) <= 2e-2 * np.max(np.abs(ref))


def test_complex_blocked_equals_single_cta_kernel(lair):
    """The blocked complex sweep against the exact in-place kernel it replaces (option cx_blocked = 0): same pivots,
    L\\U to rounding; layouts, a late zero pivot, and the solve through the factors."""
    from lair_b200 import _ffi
    rng = np.random.default_rng(77)
    a0 = _rand(rng, (300, 300), np.complex128, "normal")
    out = {}
    default = _ffi.get_option("cx_blocked")
    try:
        for v in (0, 2, 1):
            _ffi.set_option("cx_blocked", v)
            a = a0.copy()
            out[v] = (lair.lapack.getrf(a), a)
            f = np.asfortranarray(a0.copy())
            pf, sf = lair.lapack.getrf(f)
            assert pf == out[v][0][0] and sf is None
            assert np.max(np.abs(f - a)) <= 1e-9 * np.max(np.abs(a))
    finally:
        _ffi.set_option("cx_blocked", default)
    assert out[0][0] == out[1][0] == out[2][0]
    assert np.max(np.abs(out[0][1] - out[1][1])) <= 1e-10 * np.max(np.abs(out[0][1]))
    assert np.array_equal(out[1][1], out[2][1])  # cluster leaf == single-CTA leaf, bit for bit
    # zero trailing block: every step from 150 on is singular, the LAST one is reported (getrf.rs:72-73)
    z = np.zeros((300, 300), dtype=np.complex128)
    z[:150, :150] = a0[:150, :150]
    zz, zref = z.copy(), z.copy()
    piv, sing = lair.lapack.getrf(zz)
    piv_o, sing_o = oracle.getrf(zref)
    assert sing == sing_o == 299 and piv == piv_o
    # solve through the blocked factors
    b = _rand(rng, (300,), np.complex128, "normal")
    x = lair.equation.solve(a0, b)
    res = np.linalg.norm(a0 @ x - b) / (np.linalg.norm(a0) * np.linalg.norm(x) * 300 * np.finfo(np.float64).eps)
    assert res < 1.0, res


@pytest.mark.parametrize("dt", [np.complex128, np.complex64])
@pytest.mark.parametrize("n,nrhs", [(129, 1), (300, 5), (1000, 64), (777, 3)])
def test_complex_blocked_getrs_matches_oracle(lair, dt, n, nrhs):
    """Blocked complex solve (laswp on the real view, 32-row triangles + packed real GEMM): residual within 10x the
    oracle's getrs on the same factors, per column; 1-D b equals column 0 of the 2-D call."""
    rng = np.random.default_rng(n * 5 + nrhs)
    a0 = _rand(rng, (n, n), dt, "normal")
    b = _rand(rng, (n, nrhs), dt, "normal")
    lu = a0.copy()
    piv, _ = oracle.getrf(lu)
    x = lair.lapack.getrs(lu, piv, b)
    assert x.shape == (n, nrhs)
    eps = np.finfo(dt).eps / 2
    a128, b128 = a0.astype(np.complex128), b.astype(np.complex128)
    for r in range(0, nrhs, max(1, nrhs // 4)):
        xo = oracle.getrs(lu, piv, np.ascontiguousarray(b[:, r]))
        res = np.linalg.norm(a128 @ x[:, r].astype(np.complex128) - b128[:, r]) / (np.linalg.norm(a128) * np.linalg.norm(x[:, r]) * n * eps)
        res_o = np.linalg.norm(a128 @ xo.astype(np.complex128) - b128[:, r]) / (np.linalg.norm(a128) * np.linalg.norm(xo) * n * eps)
        assert res <= 10 * max(res_o, 0.01), (res, res_o)
        assert np.max(np.abs(x[:, r] - xo)) <= 1e4 * eps * n * np.max(np.abs(xo))
    one = lair.lapack.getrs(lu, piv, np.ascontiguousarray(b[:, 0]))
    assert one.shape == (n,) and np.allclose(one, x[:, 0], rtol=0, atol=1e3 * np.finfo(dt).eps * np.max(np.abs(x)))


@pytest.mark.parametrize("dt", [np.complex128, np.complex64])
def test_complex_blocked_lookahead_identical(lair, dt):
    """The complex sweep with the next block's panel recursion overlapped with the trailing update (lookahead) is the
    same arithmetic in a different schedule: identical bits, square / tall / wide."""
    from lair_b200 import _ffi
    rng = np.random.default_rng(2024)
    for shape in ((700, 700), (900, 400), (400, 900)):
        a0 = _rand(rng, shape, dt, "normal")
        res = {}
        try:
            for look in (1, 0):
                _ffi.set_option("lookahead", look)
                a = a0.copy()
                res[look] = (lair.lapack.getrf(a), a)
        finally:
            _ffi.set_option("lookahead", 1)
        assert res[0][0] == res[1][0]
        assert np.array_equal(res[0][1], res[1][1])


def test_shutdown_then_init_rebinds_cleanly(lair):
    """lair_b200_shutdown() + lair_b200_init(): kernel attributes cached in function-local statics (dynamic shared memory
    above 48 KB, cluster sizes, occupancy) and module-owned device buffers (laswp_perm, trsm_ll) must be redone, not
    reused -- the same calls give the same bits before and after the rebind."""
    from lair_b200 import _ffi
    L = _ffi.lib()
    rng = np.random.default_rng(77)
    a0 = _rand(rng, (700, 700), np.float64)
    b = _rand(rng, (700, 40), np.float64)
    m0 = _rand(rng, (300, 32, 32), np.float32)

    def run():
        a = a0.copy()
        piv, sing = lair.lapack.getrf(a)
        x = lair.lapack.getrs(a, piv, b)          # dataflow solves + collapsed permutation: module-owned buffers
        m = m0.copy()
        ip, info = lair.lapack.getrf_batched(m)
        return piv, a, x, m, ip

    first = run()
    _ffi.check(L.lair_b200_shutdown())
    _ffi.check(L.lair_b200_init(0))
    second = run()
    assert first[0] == second[0]
    for u, v in zip(first[1:], second[1:]):
        assert np.array_equal(u, v)


@pytest.mark.parametrize("shape", [(256, 256, 32), (1000, 900, 96), (2048, 2044, 128), (4096, 3072, 256)])
def test_sgemm_tf32x3_tensor_path_matches_f64_product(lair, shape):
    """f32 C -= A B on the tcgen05 tensor cores (3xTF32 split, gemm_tf32.cu) against the product in f64: the error stays
    at the level of an f32 dot product of length k (the FP32 FMA kernel's own error on the same data is the yardstick),
    tiles that overhang m / n are handled, columns beyond n are not touched.  Contraction: src/blas/gemm.rs:6-32."""
    import torch
    from lair_b200 import _ffi
    m, n, k = shape
    pad = 4
    L = _ffi.lib()
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda")
    g.manual_seed(m + 3 * n + 7 * k)
    a = torch.rand(m, k, dtype=torch.float32, device="cuda", generator=g) * 2 - 1
    b = torch.rand(k, n + pad, dtype=torch.float32, device="cuda", generator=g) * 2 - 1
    c0 = torch.rand(m, n + pad, dtype=torch.float32, device="cuda", generator=g) * 10
    ref = c0[:, :n].double() - a.double() @ b[:, :n].double()
    errs = {}
    try:
        for mode in (1, 0):
            _ffi.set_option("sgemm_tf32", mode)
            c = c0.clone()
            _ffi.check(L.lair_b200_sgemm_minus_dev(m, n, k, a.data_ptr(), k, b.data_ptr(), n + pad, c.data_ptr(), n + pad, st))
            torch.cuda.synchronize()
            assert torch.equal(c[:, n:], c0[:, n:]), "columns beyond n were written"
            errs[mode] = float((c[:, :n].double() - ref).abs().max())
    finally:
        _ffi.set_option("sgemm_tf32", 1)
    unit = 10 * (k ** 0.5) * 2.0 ** -24          # ~ the rounding of an f32 dot product of length k on data in [-1, 1]
    assert errs[0] <= 4 * unit, errs
    assert errs[1] <= 6 * unit and errs[1] <= 2 * errs[0] + unit, errs
