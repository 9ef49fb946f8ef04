"""One launch of the f32 tensor-core update at 16256 x 16256 x 128 (for an ncu capture)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lair_b200 import _ffi
L = _ffi.lib(); _ffi.check(L.lair_b200_init(0))
m = n = 16256; k = int(os.environ.get("K", "128"))
a = torch.rand(m, k, dtype=torch.float32, device="cuda"); b = torch.rand(k, n, dtype=torch.float32, device="cuda"); c = torch.rand(m, n, dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    _ffi.check(L.lair_b200_sgemm_minus_dev(m, n, k, a.data_ptr(), k, b.data_ptr(), n, c.data_ptr(), n, st))
torch.cuda.synchronize()
print("done")
