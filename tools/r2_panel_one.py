"""One panel launch per shape for ncu: python tools/r2_panel_one.py [d|s] m w"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lair_b200 import _ffi
L = _ffi.lib()
pfx, m, w = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
dt = torch.float64 if pfx == "d" else torch.float32
a0 = torch.rand(m, w, dtype=dt, device="cuda")
ipiv = torch.empty(w, dtype=torch.int32, device="cuda")
info = torch.empty(1, dtype=torch.int32, device="cuda")
fn = getattr(L, f"lair_b200_{pfx}getrf_dev")
for _ in range(3):
    a = a0.clone()
    _ffi.check(fn(m, w, a.data_ptr(), w, ipiv.data_ptr(), info.data_ptr(), torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
print("ok", int(info.item()))
