"""128- / 64-wide block steps sharing one trailing GEMM (pair_small): bit-identity and time at n = 4096 ... 16384."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lair_b200 import _ffi
L = _ffi.lib(); _ffi.check(L.lair_b200_init(0))
st = torch.cuda.current_stream().cuda_stream
def ev_time(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)
for pfx, dt, opt in (("d", torch.float64, "pair_small"), ("s", torch.float32, "pair_small_f32")):
    fn = getattr(L, f"lair_b200_{pfx}getrf_dev")
    it = torch.int64 if pfx == "d" else torch.int32
    for n in (4096, 8192, 16384):
        a0 = torch.rand(n, n, dtype=dt, device="cuda") * 10
        a = torch.empty_like(a0)
        ipiv = torch.empty(n, dtype=torch.int32, device="cuda"); info = torch.empty(1, dtype=torch.int32, device="cuda")
        ref = None
        for v in (0, 512, 1536, 3072, 6144):
            _ffi.set_option(opt, v)
            ts = []
            for rep in range(4):
                a.copy_(a0)
                ts.append(ev_time(lambda: _ffi.check(fn(n, n, a.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), st))))
            if ref is None:
                ref = (a.clone(), ipiv.clone()); same = True
            else:
                same = bool(torch.equal(ref[1], ipiv) and torch.equal(ref[0].view(it), a.view(it)))
            print(json.dumps({"bench": f"{pfx}getrf_pair_small", "n": n, opt: v, "ms": round(min(ts[1:]), 3), "tflops": round(2 / 3 * n ** 3 / min(ts[1:]) * 1e-9, 2), "same_bits": same}), flush=True)
        _ffi.set_option(opt, 0)
        del a0, a, ref
