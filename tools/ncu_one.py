"""One device-resident call for an ncu capture: python tools/ncu_one.py {dgeqrf|sgeqrf|zgetrf|cgetrf} n"""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lair_b200 import _ffi  # noqa: E402

what, n = sys.argv[1], int(sys.argv[2])
L = _ffi.lib()
s = torch.cuda.current_stream().cuda_stream
pfx = what[0]
rdt = torch.float64 if pfx in "dz" else torch.float32
if what.endswith("geqrf"):
    a = torch.rand(n, n, dtype=rdt, device="cuda") * 10
    tau = torch.empty(n, dtype=rdt, device="cuda")
    _ffi.check(getattr(L, f"lair_b200_{pfx}geqrf_dev")(n, n, a.data_ptr(), n, tau.data_ptr(), s))
else:
    a = torch.complex(torch.rand(n, n, dtype=rdt, device="cuda") * 10, torch.rand(n, n, dtype=rdt, device="cuda") * 10)
    ipiv = torch.empty(n, dtype=torch.int32, device="cuda")
    info = torch.empty(1, dtype=torch.int32, device="cuda")
    _ffi.check(getattr(L, f"lair_b200_{pfx}getrf_dev")(n, n, a.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), s))
torch.cuda.synchronize()
print("done", what, n)
