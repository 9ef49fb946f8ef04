"""GPU, >= 2 devices: the distributed block-cyclic LU (mg.cu) against the single-GPU path, as a test the driver runs.

Launches tools/mg_check.py under torchrun on 2 ranks (NCCL over NVLink): for n = 1024 ... 4096 (f64, f32, several block
widths, a ragged size) the pivots must be IDENTICAL to lair_b200_dgetrf_dev's and L\\U within 1e-9 max|LU| (f64: the two
paths subtract l*u with an FMA in the same ascending-k order).  The single-GPU path itself is tied to the oracle by
tests/test_gpu_parity.py.  Skips on a one-GPU box.  The index arithmetic of the distribution is covered on CPU by
tests/test_sharding_gloo.py.
"""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count() -> int:
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2])
def test_getrf_mg_matches_single_gpu(world):
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {_gpu_count()}")
    port = 29500 + (os.getpid() % 400)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "mg_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    lines = [json.loads(ln) for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert len(lines) >= 5 and all(ln["ok"] for ln in lines), lines
    assert all(ln["pivots_identical"] for ln in lines if "float64" in ln["dtype"]), lines


def test_batched_mg_entry_one_gpu_and_too_many(tmp_path):
    """lair_b200_*getrf_batched_mg with ngpu = 1 is the single-GPU call; more GPUs than visible is refused loudly."""
    import numpy as np
    import lair_b200
    from lair_b200 import _ffi
    import oracle
    rng = np.random.default_rng(8)
    for dt in (np.float32, np.float64):
        a0 = rng.uniform(0, 10, size=(3001, 32, 32)).astype(dt)
        ref = a0.copy()
        piv_o, info_o = oracle.getrf_batched(ref)
        a = a0.copy()
        fn = getattr(_ffi.lib(), "lair_b200_%sgetrf_batched_mg" % ("s" if dt == np.float32 else "d"))
        ipiv = np.zeros((3001, 32), dtype=np.int32)
        info = np.zeros(3001, dtype=np.int32)
        _ffi.check(fn(3001, 32, a.ctypes.data, ipiv.ctypes.data, info.ctypes.data, 1))
        assert np.array_equal(a, ref) and np.array_equal(ipiv, piv_o.astype(np.int32)) and np.array_equal(info, info_o.astype(np.int32))
    with pytest.raises(_ffi.LairB200Error):
        lair_b200.lapack.getrf_batched(np.zeros((4, 32, 32)), ngpu=_gpu_count() + 1)


@pytest.mark.parametrize("ngpu", [2, 4, 8])
def test_batched_mg_entry_matches_oracle(ngpu):
    """One process, `ngpu` devices, contiguous slices of the batch: bit-identical to the oracle (ragged batch, both types,
    n = 32 and a small n), i.e. to the single-GPU entry."""
    if _gpu_count() < ngpu:
        pytest.skip(f"needs {ngpu} GPUs, this box has {_gpu_count()}")
    import numpy as np
    import lair_b200
    import oracle
    rng = np.random.default_rng(80 + ngpu)
    for dt, n, batch in ((np.float32, 32, 40_003), (np.float64, 32, 20_001), (np.float64, 7, 1_003)):
        a0 = rng.uniform(0, 10, size=(batch, n, n)).astype(dt)
        a0[5] = 0                      # a singular matrix in the first slice
        a0[batch - 2, :, 1] = 0        # and one in the last
        ref = a0.copy()
        piv_o, info_o = oracle.getrf_batched(ref)
        a = a0.copy()
        ipiv, info = lair_b200.lapack.getrf_batched(a, ngpu=ngpu)
        assert np.array_equal(a, ref, equal_nan=True)
        assert np.array_equal(ipiv, piv_o.astype(np.int32)) and np.array_equal(info, info_o.astype(np.int32))
