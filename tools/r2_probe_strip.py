"""trsm_strip (one register-tiled triangle launch for the wide trailing ranges) against the chain of fused launches:
bit-identity of pivots and L\\U, time, per-family kernel time."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lair_b200 import _ffi
L = _ffi.lib()
def stream(): return torch.cuda.current_stream().cuda_stream
def bits(x): return x.view(torch.int64 if x.dtype == torch.float64 else torch.int32)
for pfx, dt, m, n in (("d", torch.float64, 3000, 3000), ("s", torch.float32, 5000, 3100), ("s", torch.float32, 16384, 16384), ("s", torch.float32, 8192, 8192),
                      ("d", torch.float64, 4096, 4096), ("d", torch.float64, 8192, 8192), ("d", torch.float64, 16384, 16384)):
    fn = getattr(L, f"lair_b200_{pfx}getrf_dev")
    a0 = torch.rand(m, n, dtype=dt, device="cuda") * 10
    a = a0.clone()
    k = min(m, n)
    ipiv = torch.empty(k, dtype=torch.int32, device="cuda"); info = torch.empty(1, dtype=torch.int32, device="cuda")
    ref = None
    for strip in (0, 1, 2):
        _ffi.set_option("trsm_strip", strip)
        ts = []
        for rep in range(5):
            a.copy_(a0); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); _ffi.check(fn(m, n, a.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), stream())); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res = (a.clone(), ipiv.clone())
        if ref is None: ref = res
        same = bool(torch.equal(bits(res[0]), bits(ref[0])) and torch.equal(res[1], ref[1]))
        a.copy_(a0); _ffi.profile_begin(); _ffi.check(fn(m, n, a.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), stream())); torch.cuda.synchronize(); fam = _ffi.profile_end()
        flops = m * n * n - n ** 3 / 3 if m >= n else 0
        print(json.dumps({"bench": f"{pfx}getrf_strip", "m": m, "n": n, "trsm_strip": strip, "ms_best": round(min(ts[1:]), 3), "tflops": round(flops / min(ts[1:]) * 1e-9, 2),
                          "same_bits_as_strip0": same, "fam_ms": {k_: round(v["ms"], 2) for k_, v in fam.items() if v["launches"]},
                          "fam_launches": {k_: v["launches"] for k_, v in fam.items() if v["launches"]}}), flush=True)
    del ref
    _ffi.set_option("trsm_strip", 1)
