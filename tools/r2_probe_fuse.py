import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lair_b200 import _ffi
L = _ffi.lib()
def stream(): return torch.cuda.current_stream().cuda_stream
for pfx, dt, n in (("s", torch.float32, 16384), ("s", torch.float32, 8192), ("d", torch.float64, 8192), ("d", torch.float64, 16384)):
    fn = getattr(L, f"lair_b200_{pfx}getrf_dev")
    a0 = torch.rand(n, n, dtype=dt, device="cuda") * 10
    a = a0.clone()
    ipiv = torch.empty(n, dtype=torch.int32, device="cuda"); info = torch.empty(1, dtype=torch.int32, device="cuda")
    for fuse in (1, 2, 0):
        _ffi.set_option("fuse_swap_trsm", fuse)
        ts = []
        for rep in range(5):
            a.copy_(a0); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); _ffi.check(fn(n, n, a.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), stream())); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        a.copy_(a0); _ffi.profile_begin(); _ffi.check(fn(n, n, a.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), stream())); torch.cuda.synchronize(); fam = _ffi.profile_end()
        print(json.dumps({"bench": f"{pfx}getrf_fuse", "n": n, "fuse": fuse, "ms_best": round(min(ts[1:]), 3), "tflops": round(2 / 3 * n ** 3 / min(ts[1:]) * 1e-9, 2),
                          "fam_ms": {k: round(v["ms"], 2) for k, v in fam.items() if v["launches"]}, "fam_launches": {k: v["launches"] for k, v in fam.items() if v["launches"]}}), flush=True)
    _ffi.set_option("fuse_swap_trsm", 1)
