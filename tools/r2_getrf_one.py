"""One device-resident getrf for ncu: python tools/r2_getrf_one.py [d|s] n"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lair_b200 import _ffi
L = _ffi.lib()
pfx, n = sys.argv[1], int(sys.argv[2])
dt = torch.float64 if pfx == "d" else torch.float32
a0 = torch.rand(n, n, dtype=dt, device="cuda") * 10
ipiv = torch.empty(n, dtype=torch.int32, device="cuda"); info = torch.empty(1, dtype=torch.int32, device="cuda")
fn = getattr(L, f"lair_b200_{pfx}getrf_dev")
rt = torch.cuda.cudart()
for rep in range(2):
    a = a0.clone(); torch.cuda.synchronize()
    if rep == 1: rt.cudaProfilerStart()
    _ffi.check(fn(n, n, a.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    if rep == 1: rt.cudaProfilerStop()
print("ok", int(info.item()))
