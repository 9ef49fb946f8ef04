"""f64 GEMM CTA order (gemm_raster): rank-256 update at the shapes of the n = 65 536 factorization."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lair_b200 import _ffi
L = _ffi.lib(); _ffi.check(L.lair_b200_init(0))
st = torch.cuda.current_stream().cuda_stream
def run(m, n, k, raster):
    _ffi.set_option("gemm_raster", raster)
    a = torch.rand(m, k, dtype=torch.float64, device="cuda"); b = torch.rand(k, n, dtype=torch.float64, device="cuda")
    c = torch.rand(m, n, dtype=torch.float64, device="cuda")
    ts = []
    for i in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); _ffi.check(L.lair_b200_dgemm_minus_dev(m, n, k, a.data_ptr(), k, b.data_ptr(), n, c.data_ptr(), n, st)); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = min(ts[1:])
    print(json.dumps({"bench": "dgemm_minus", "m": m, "n": n, "k": k, "raster": raster, "ms": ms, "tflops": 2.0 * m * n * k / ms * 1e-9}), flush=True)
    del a, b, c
for (m, n, k) in ((65280, 8192, 256), (65280, 32768, 256), (32768, 8192, 256), (8192, 8192, 256), (7808, 7680, 128), (32768, 32768, 256)):
    for r in (1, 4, 8, 16):
        run(m, n, k, r)
# check: raster must not change the result
m, n, k = 3000, 2100, 256
a = torch.rand(m, k, dtype=torch.float64, device="cuda"); b = torch.rand(k, n, dtype=torch.float64, device="cuda"); c0 = torch.rand(m, n, dtype=torch.float64, device="cuda")
outs = []
for r in (1, 8, 5):
    _ffi.set_option("gemm_raster", r); c = c0.clone()
    _ffi.check(L.lair_b200_dgemm_minus_dev(m, n, k, a.data_ptr(), k, b.data_ptr(), n, c.data_ptr(), n, st)); torch.cuda.synchronize(); outs.append(c)
print(json.dumps({"check": "raster_invariance", "identical": bool(torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])),
                  "max_err_vs_torch": float((outs[1] - (c0 - a @ b)).abs().max())}))
