"""Host-pointer getrf of n = 65 536 from / to a pinned array (what bench.py's e2e leg times), pairing off / on."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lair_b200
from lair_b200 import _ffi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
host = torch.empty(n, n, dtype=torch.float64).pin_memory()
src = torch.rand(n, n, dtype=torch.float64, device="cuda")
h = host.numpy()
res = {}
div_list = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["4"])]
for pair, div in [(p_, d_) for d_ in div_list for p_ in ((0, 16384, 16384) if len(div_list) == 1 else (16384, 16384))]:
    _ffi.set_option("pair_k512", pair)
    _ffi.set_option("stream_join_div", div)
    host.copy_(src); torch.cuda.synchronize()
    t0 = time.perf_counter(); piv, sing = lair_b200.lapack.getrf(h); t1 = time.perf_counter()
    chk = float(h[::4097, ::4099].sum()) + float(h[-1, -1])
    print(json.dumps({"bench": "dgetrf_host_pinned", "n": n, "pair_k512": pair, "stream_join_div": div, "ms": round((t1 - t0) * 1e3, 1), "tflops": round(2 / 3 * n ** 3 / (t1 - t0) * 1e-12, 2),
                      "piv_sum": int(np.sum(piv)), "sample_checksum": chk, "sing": sing}), flush=True)
