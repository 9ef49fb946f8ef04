"""Round-2 first contact: host memory of the box, n = 65 536 f64 getrf on ONE GPU (device-resident),
the current batched kernel (the baseline this round starts from)."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lair_b200 import _ffi
L = _ffi.lib()
_ffi.check(L.lair_b200_init(0))
def out(**kw): print(json.dumps(kw), flush=True)
import psutil
vm = psutil.virtual_memory()
out(host_total_gib=vm.total / 2**30, host_avail_gib=vm.available / 2**30, cpus=os.cpu_count(),
    cgroup_max=open("/sys/fs/cgroup/memory.max").read().strip() if os.path.exists("/sys/fs/cgroup/memory.max") else None)
stream = torch.cuda.current_stream().cuda_stream
for n in (16384, 32768, 65536):
    a0 = torch.rand(n, n, dtype=torch.float64, device="cuda") * 10
    a = torch.empty_like(a0)
    ipiv = torch.empty(n, dtype=torch.int32, device="cuda")
    info = torch.empty(1, dtype=torch.int32, device="cuda")
    ts = []
    for rep in range(3 if n < 65536 else 2):
        a.copy_(a0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _ffi.check(L.lair_b200_dgetrf_dev(n, n, a.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), stream))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    _ffi.check_fault(stream)
    out(bench="dgetrf_dev_1gpu", n=n, ms=ts, tflops=2 / 3 * n ** 3 / min(ts) * 1e-9, info=int(info.item()),
        mem_gib=torch.cuda.max_memory_allocated() / 2**30)
    del a0, a
    torch.cuda.empty_cache()
