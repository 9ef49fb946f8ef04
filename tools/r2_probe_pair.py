"""Two block steps sharing one K = 512 trailing GEMM (pair_k512) against the step-by-step sweep at large n."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lair_b200 import _ffi
L = _ffi.lib(); _ffi.check(L.lair_b200_init(0))
st = torch.cuda.current_stream().cuda_stream
def ev_time(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)
for n in (24576, 32768, 65536):
    a0 = torch.rand(n, n, dtype=torch.float64, device="cuda")
    a = torch.empty_like(a0)
    ipiv = torch.empty(n, dtype=torch.int32, device="cuda"); info = torch.empty(1, dtype=torch.int32, device="cuda")
    ref = None
    for pair in (0, 16384, 8192):
        _ffi.set_option("pair_k512", pair)
        ts = []
        for rep in range(2):
            a.copy_(a0)
            ts.append(ev_time(lambda: _ffi.check(L.lair_b200_dgetrf_dev(n, n, a.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), st))))
        if ref is None:
            ref = (a[:, : min(n, 8192)].clone(), ipiv.clone())   # (a column slab is enough to see any divergence; 34 GB copies do not fit twice)
            same = True
        else:
            same = bool(torch.equal(ref[1], ipiv) and torch.equal(ref[0].view(torch.int64), a[:, : min(n, 8192)].view(torch.int64)))
        print(json.dumps({"bench": "dgetrf_pair", "n": n, "pair_k512": pair, "ms": round(min(ts), 1), "tflops": round(2 / 3 * n ** 3 / min(ts) * 1e-9, 2),
                          "same_pivots_and_bits_as_unpaired": same, "info": int(info.item())}), flush=True)
    _ffi.set_option("pair_k512", 16384)
    del a0, a, ref
