"""How fast is cudaHostRegister / cudaHostUnregister on this box, against a pageable cudaMemcpy?"""
import json, time
import numpy as np, torch
rt = torch.cuda.cudart()
torch.cuda.init()
for mb in (64, 512):
    n = mb << 20
    a = np.empty(n, dtype=np.uint8); a[:] = 1   # touched pages
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for rep in range(3):
        t0 = time.perf_counter(); r = rt.cudaHostRegister(a.ctypes.data, n, 0); t1 = time.perf_counter()
        t = torch.from_numpy(a)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        d.copy_(t, non_blocking=True); torch.cuda.synchronize(); t3 = time.perf_counter()
        u = rt.cudaHostUnregister(a.ctypes.data); t4 = time.perf_counter()
        d.copy_(t); torch.cuda.synchronize(); t5 = time.perf_counter()
        print(json.dumps({"mb": mb, "rep": rep, "register_ms": (t1 - t0) * 1e3, "h2d_registered_ms": (t3 - t2) * 1e3, "unregister_ms": (t4 - t3) * 1e3,
                          "h2d_pageable_ms": (t5 - t4) * 1e3, "rc": [int(r), int(u)]}), flush=True)
    # chunked registration: 16 MB pieces
    t0 = time.perf_counter()
    step = 16 << 20
    for o in range(0, n, step): rt.cudaHostRegister(a.ctypes.data + o, min(step, n - o), 0)
    t1 = time.perf_counter()
    for o in range(0, n, step): rt.cudaHostUnregister(a.ctypes.data + o)
    t2 = time.perf_counter()
    print(json.dumps({"mb": mb, "chunked_16mb_register_ms": (t1 - t0) * 1e3, "chunked_unregister_ms": (t2 - t1) * 1e3}), flush=True)
    # memcpy speed into a pinned bounce buffer (what a staging thread would do)
    p = torch.empty(n, dtype=torch.uint8, pin_memory=True).numpy()
    t0 = time.perf_counter(); np.copyto(p, a); t1 = time.perf_counter()
    print(json.dumps({"mb": mb, "memcpy_to_pinned_ms": (t1 - t0) * 1e3, "gbps": n / (t1 - t0) / 1e9}), flush=True)
