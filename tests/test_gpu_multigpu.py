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
