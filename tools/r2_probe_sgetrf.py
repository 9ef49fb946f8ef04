"""sgetrf with the tensor-core (3xTF32) trailing update on / off: time and backward error."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lair_b200 import _ffi
import devcheck
L = _ffi.lib(); _ffi.check(L.lair_b200_init(0))
st = torch.cuda.current_stream().cuda_stream
for (m, n) in ((4096, 4096), (8192, 8192), (16384, 16384), (262144, 1024)):
    g = torch.Generator(device="cuda"); g.manual_seed(m + n)
    a0 = torch.rand(m, n, dtype=torch.float32, device="cuda", generator=g) * 10
    for mode, nb in ((0, 0), (1, 0), (1, 128), (1, 256), (1, 512)):
        _ffi.set_option("sgemm_tf32", mode); _ffi.set_option("nb", nb)
        a = a0.clone(); ipiv = torch.empty(min(m, n), dtype=torch.int32, device="cuda"); info = torch.empty(1, dtype=torch.int32, device="cuda")
        ts = []
        for r in range(3):
            a.copy_(a0); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); _ffi.check(L.lair_b200_sgetrf_dev(m, n, a.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), st)); e1.record()
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        _ffi.check_fault(st)
        be = devcheck.backward_error_dev(a0, a, ipiv)
        flops = m * n * n - n ** 3 / 3.0
        print(json.dumps({"bench": "sgetrf", "m": m, "n": n, "tf32x3": mode, "nb": nb, "ms": min(ts), "tflops": flops / min(ts) * 1e-9, "backward_error": be, "info": int(info.item())}), flush=True)
    _ffi.set_option("nb", 0)
