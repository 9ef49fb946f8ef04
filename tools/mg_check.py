"""Multi-GPU parity check, run under torchrun on >= 2 GPUs of one node:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mg_check.py

The distributed block-cyclic LU must give the pivots and L\\U of the single-GPU path (which the
parity tests tie to the oracle): identical pivots, |LU - LU_1gpu| <= 1e-9 max|LU|.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lair_b200  # noqa: E402
from lair_b200 import _ffi, multigpu  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    multigpu.init()
    ok = True
    for n, nb, dt in ((1024, 128, torch.float64), (2048, 256, torch.float64), (3000, 128, torch.float64), (4096, 512, torch.float64),
                      (2048, 128, torch.float32)):
        gen = torch.Generator(device="cuda")
        gen.manual_seed(1234 + n)
        a_full = torch.rand(n, n, dtype=dt, device="cuda", generator=gen) * 10
        a_loc = multigpu.distribute_columns(a_full, nb)
        ipiv, info = multigpu.getrf_mg(a_loc, n, nb)
        torch.cuda.synchronize()
        lu = multigpu.gather_columns(a_loc, n, nb)
        # single-GPU reference through the same C ABI
        ref = a_full.clone()
        ipiv1 = torch.empty(n, dtype=torch.int32, device="cuda")
        info1 = torch.empty(1, dtype=torch.int32, device="cuda")
        pfx = "d" if dt == torch.float64 else "s"
        _ffi.check(getattr(_ffi.lib(), f"lair_b200_{pfx}getrf_dev")(n, n, ref.data_ptr(), n, ipiv1.data_ptr(), info1.data_ptr(),
                                                                     torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        same_piv = bool(torch.equal(ipiv, ipiv1))
        err = float((lu - ref).abs().max() / ref.abs().max())
        tol = 1e-9 if dt == torch.float64 else 1e-3
        good = (same_piv or dt == torch.float32) and err <= tol and int(info.item()) == int(info1.item())
        ok &= good
        if rank == 0:
            print(json.dumps({"check": "getrf_mg_vs_1gpu", "n": n, "nb": nb, "dtype": str(dt), "world": world, "pivots_identical": same_piv,
                              "max_rel_diff": err, "info": int(info.item()), "ok": good}), flush=True)
    multigpu.finalize()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
