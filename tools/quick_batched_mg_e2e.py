"""End-to-end mats/s of the single-process multi-GPU batched entry (pinned host buffers): ngpu = 1, 2, 4, 8."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lair_b200
from lair_b200 import _ffi
ng = torch.cuda.device_count()
for dt, tdt in ((np.float32, torch.float32), (np.float64, torch.float64)):
    batch = 800_000
    src = torch.empty(batch, 32, 32, dtype=tdt).pin_memory()
    src.uniform_(0, 10)
    work = torch.empty_like(src).pin_memory()
    w = work.numpy()
    ref = None
    for g in [x for x in (1, 2, 4, 8) if x <= ng]:
        ts = []
        for rep in range(3):
            work.copy_(src)
            t0 = time.perf_counter()
            ipiv, info = lair_b200.lapack.getrf_batched(w, ngpu=g)
            ts.append(time.perf_counter() - t0)
        if ref is None:
            ref = (w.copy(), ipiv.copy())
        same = bool(np.array_equal(ref[0], w) and np.array_equal(ref[1], ipiv))
        t = min(ts[1:])
        print(json.dumps({"bench": "getrf_batched_mg_e2e", "dtype": dt.__name__, "batch": batch, "ngpu": g, "ms": t * 1e3, "mats_per_s": batch / t,
                          "host_gbs_each_way": batch * 32 * 32 * np.dtype(dt).itemsize / t * 1e-9, "identical_to_1gpu": same}), flush=True)
