"""Per-family kernel time (event profiler) of one getrf: python tools/r2_probe_families.py"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lair_b200 import _ffi
L = _ffi.lib()
for pfx, dt, n in (("s", torch.float32, 16384), ("s", torch.float32, 8192), ("d", torch.float64, 8192), ("d", torch.float64, 4096)):
    fn = getattr(L, f"lair_b200_{pfx}getrf_dev")
    a0 = torch.rand(n, n, dtype=dt, device="cuda") * 10
    a = a0.clone()
    ipiv = torch.empty(n, dtype=torch.int32, device="cuda"); info = torch.empty(1, dtype=torch.int32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        a.copy_(a0); _ffi.check(fn(n, n, a.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), s))
    torch.cuda.synchronize()
    a.copy_(a0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); _ffi.check(fn(n, n, a.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), s)); e1.record(); torch.cuda.synchronize()
    plain = e0.elapsed_time(e1)
    a.copy_(a0)
    _ffi.profile_begin()
    _ffi.check(fn(n, n, a.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), s))
    torch.cuda.synchronize()
    fam = _ffi.profile_end()
    print(json.dumps({"bench": f"{pfx}getrf_families", "n": n, "ms_unprofiled": round(plain, 3),
                      "families": {k: {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items()} if isinstance(v, dict) else v for k, v in fam.items()}}), flush=True)
