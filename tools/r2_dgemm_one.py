"""One launch of the DMMA trailing update at the shape of the first block step of n = 65 536 (65280 x 65280 x 256, C 34 GB, in
place at leading dimension 65536) for an ncu capture: python tools/r2_dgemm_one.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lair_b200 import _ffi
L = _ffi.lib(); _ffi.check(L.lair_b200_init(0))
ld = 65536; m = n = ld - 256; k = 256
mat = torch.rand(ld, ld, dtype=torch.float64, device="cuda")        # the whole matrix: A = L21 (below the block), B = U12 (right of it)
a_ptr = mat.data_ptr() + (256 * ld) * 8
b_ptr = mat.data_ptr() + 256 * 8
c_ptr = mat.data_ptr() + (256 * ld + 256) * 8
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    _ffi.check(L.lair_b200_dgemm_minus_dev(m, n, k, a_ptr, ld, b_ptr, ld, c_ptr, ld, st))
torch.cuda.synchronize()
print("done")
