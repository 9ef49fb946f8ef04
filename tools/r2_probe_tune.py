"""Sweep thresholds of the blocked sweep again with the fourth-generation panel kernel: dgetrf n = 8192 and sgetrf n = 16384."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lair_b200 import _ffi
L = _ffi.lib()
def stream(): return torch.cuda.current_stream().cuda_stream
def timeit(fn, setup, reps=4, warm=1):
    for _ in range(warm): setup(); fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        setup(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))
for pfx, dt, n, t1s, t2s in (("d", torch.float64, 8192, (3072, 4096, 5120, 6144, 7168, 8192), (10240,)),
                             ("s", torch.float32, 16384, (4096, 6144, 8192, 10240, 12288), (10240, 12288, 14336, 16384))):
    fn = getattr(L, f"lair_b200_{pfx}getrf_dev")
    a0 = torch.rand(n, n, dtype=dt, device="cuda") * 10
    a = a0.clone()
    ipiv = torch.empty(n, dtype=torch.int32, device="cuda"); info = torch.empty(1, dtype=torch.int32, device="cuda")
    d_t2 = _ffi.get_option("nb_t2")
    for t2 in t2s:
        for t1 in t1s:
            if t1 > t2: continue
            for cop in (0, 2048, 3072, 4096, 6144, 1 << 30):
                _ffi.set_option("nb_t1", t1); _ffi.set_option("nb_t2", t2); _ffi.set_option("chain_on_p", cop)
                best, med = timeit(lambda: _ffi.check(fn(n, n, a.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), stream())), setup=lambda: a.copy_(a0))
                print(json.dumps({"bench": f"{pfx}getrf_tune", "n": n, "nb_t1": t1, "nb_t2": t2, "chain_on_p": cop, "ms_best": round(best, 3), "ms_med": round(med, 3)}), flush=True)
    _ffi.set_option("nb_t1", 0); _ffi.set_option("nb_t2", d_t2); _ffi.set_option("chain_on_p", 3072)
