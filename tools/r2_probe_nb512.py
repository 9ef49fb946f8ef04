"""Large n: forced block width 256 vs 512 (deeper K for the DMMA update) and the GEMM alone at K = 256 / 512 / raster 8 / 32."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lair_b200 import _ffi
L = _ffi.lib(); _ffi.check(L.lair_b200_init(0))
st = torch.cuda.current_stream().cuda_stream
def ev_time(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)
# GEMM alone on the in-place layout of n = 65536
ld = 65536
mat = torch.rand(ld, ld, dtype=torch.float64, device="cuda")
for k in (256, 512):
    for raster in (8, 32):
        _ffi.set_option("gemm_raster", raster)
        m = n = ld - k
        a_ptr = mat.data_ptr() + (k * ld) * 8; b_ptr = mat.data_ptr() + k * 8; c_ptr = mat.data_ptr() + (k * ld + k) * 8
        fn = lambda: _ffi.check(L.lair_b200_dgemm_minus_dev(m, n, k, a_ptr, ld, b_ptr, ld, c_ptr, ld, st))
        fn(); torch.cuda.synchronize()
        t = min(ev_time(fn) for _ in range(2))
        print(json.dumps({"bench": "dgemm_inplace", "m": m, "n": n, "k": k, "raster": raster, "ms": round(t, 2), "tflops": round(2.0 * m * n * k / t * 1e-9, 2)}), flush=True)
_ffi.set_option("gemm_raster", 8)
del mat
for n in (32768, 65536):
    a0 = torch.rand(n, n, dtype=torch.float64, device="cuda")
    a = torch.empty_like(a0)
    ipiv = torch.empty(n, dtype=torch.int32, device="cuda"); info = torch.empty(1, dtype=torch.int32, device="cuda")
    for nb in (0, 512):
        _ffi.set_option("nb", nb)
        ts = []
        for rep in range(2):
            a.copy_(a0)
            ts.append(ev_time(lambda: _ffi.check(L.lair_b200_dgetrf_dev(n, n, a.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), st))))
        print(json.dumps({"bench": "dgetrf_nb", "n": n, "nb": nb or "auto(256)", "ms": round(min(ts), 1), "tflops": round(2 / 3 * n ** 3 / min(ts) * 1e-9, 2), "info": int(info.item())}), flush=True)
    _ffi.set_option("nb", 0)
    del a0, a
