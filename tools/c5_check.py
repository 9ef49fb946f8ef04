"""BASELINE configs[4]: tall-skinny f32 getrf (262144 x 1024) and f64 n=16384 -- time + backward error."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lair_b200 import _ffi  # noqa: E402

L = _ffi.lib()


def run(m, n, dt, pfx, reps=3):
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5)
    a0 = torch.rand(m, n, dtype=dt, device="cuda", generator=gen) * 10
    a = a0.clone()
    ipiv = torch.empty(min(m, n), dtype=torch.int32, device="cuda")
    info = torch.empty(1, dtype=torch.int32, device="cuda")
    fn = getattr(L, f"lair_b200_{pfx}getrf_dev")
    stream = torch.cuda.current_stream().cuda_stream
    ts = []
    for _ in range(reps + 1):
        a.copy_(a0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _ffi.check(fn(m, n, a.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), stream))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = min(ts[1:])
    k = min(m, n)
    piv = ipiv.cpu().numpy()
    perm = np.arange(m)
    for i, p in enumerate(piv):
        if i != p:
            perm[i], perm[p] = perm[p], perm[i]
    PA = a0.double()[torch.from_numpy(perm).cuda()]
    LU = a.double()
    Lm = torch.tril(LU[:, :k], -1)
    Lm[torch.arange(k), torch.arange(k)] = 1.0
    U = torch.triu(LU[:k, :])
    eps = 2.0 ** -53 if dt == torch.float64 else 2.0 ** -24
    be = float(torch.linalg.norm(PA - Lm @ U) / (max(m, n) * eps * torch.linalg.norm(PA)))
    flops = m * n * n - n ** 3 / 3 if m >= n else n * m * m - m ** 3 / 3
    print(json.dumps({"check": "c5", "m": m, "n": n, "dtype": str(dt), "ms": ms, "tflops": flops / ms * 1e-9, "backward_error": be,
                      "info": int(info.item()), "pivots_in_range": bool((piv >= np.arange(k)).all() and (piv < m).all())}), flush=True)


if __name__ == "__main__":
    run(262144, 1024, torch.float32, "s")
    run(65536, 512, torch.float64, "d")
    run(16384, 16384, torch.float64, "d", reps=2)
    run(16384, 16384, torch.float32, "s", reps=2)
