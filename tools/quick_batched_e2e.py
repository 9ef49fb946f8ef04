"""Torch-free check of the host-pointer batched entry (lair_b200_{s,d}getrf_batched): parity against the oracle on a
ragged batch, then the end-to-end rate with a page-locked matrix array and pageable pivot / info arrays (what
lair_b200.lapack.getrf_batched allocates).  Usage: python tools/quick_batched_e2e.py [matrices]"""
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lair_b200  # noqa: E402
import oracle  # noqa: E402

rt = ctypes.CDLL("libcudart.so.12")
rng = np.random.default_rng(7)
for dt in (np.float64, np.float32):
    a0 = rng.uniform(0, 10, size=(40001, 32, 32)).astype(dt)  # three chunks, the last one ragged
    a, ref = a0.copy(), a0[:3000].copy()
    ipiv, info = lair_b200.lapack.getrf_batched(a)
    po, io = oracle.getrf_batched(ref)
    tail = a0[-500:].copy()
    pt, it = oracle.getrf_batched(tail)
    ok = (np.array_equal(a[:3000], ref) and np.array_equal(ipiv[:3000], po.astype(np.int32)) and np.array_equal(info[:3000], io.astype(np.int32))
          and np.array_equal(a[-500:], tail) and np.array_equal(ipiv[-500:], pt.astype(np.int32)) and np.array_equal(info[-500:], it.astype(np.int32)))
    print(json.dumps({"check": "batched host entry vs oracle", "dtype": np.dtype(dt).name, "batch": 40001, "bit_identical": bool(ok)}), flush=True)
    assert ok

nmat = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
h = rng.uniform(0, 10, size=(nmat, 32, 32))
w = h.copy()
assert rt.cudaHostRegister(ctypes.c_void_p(w.ctypes.data), ctypes.c_size_t(w.nbytes), 0) == 0
lair_b200.lapack.getrf_batched(w)
for chunk in (16384, 8192, 4096, 2048):
  lair_b200._ffi.set_option("batched_chunk", chunk)
  ts = []
  for _ in range(3):
    w[...] = h
    t0 = time.perf_counter()
    p, i_ = lair_b200.lapack.getrf_batched(w)
    ts.append(time.perf_counter() - t0)
  t = min(ts)
  print(json.dumps({"probe": "c3 e2e f64, pinned matrices, pageable pivots/info", "matrices": nmat, "batched_chunk": chunk, "ms": [round(x * 1e3, 2) for x in ts],
                    "mats_per_s": nmat / t, "h2d_GBps": w.nbytes / t * 1e-9, "d2h_GBps": (w.nbytes + p.nbytes + i_.nbytes) / t * 1e-9}), flush=True)
