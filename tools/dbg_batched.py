"""Debug helper: which matrices of the variant test differ from the oracle for a batched cfg."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lair_b200 as lair
from lair_b200 import _ffi
import oracle

def run(dt, cfg):
    rng = np.random.default_rng(100 + cfg)
    a0 = ((rng.random((1001, 32, 32)) - 0.5) * 4).astype(dt)
    a0[5] = rng.integers(-2, 3, size=(32, 32)).astype(dt)
    a0[6] = 0
    a0[7, :, 3] = 0
    a0[8, 4, 4] = np.nan
    a0[9, 2, 2] = np.inf
    a0[10] = a0[10] * dt(1e-39 if dt == np.float32 else 1e-309)
    a0[11] = a0[11] * dt(1e37 if dt == np.float32 else 1e307)
    a0[1000] = 1
    ref = a0.copy()
    piv_o, info_o = oracle.getrf_batched(ref)
    _ffi.set_option("batched_cfg", cfg)
    a = a0.copy()
    ipiv, info = lair.lapack.getrf_batched(a)
    _ffi.set_option("batched_cfg", -1)
    bad = [i for i in range(a.shape[0]) if not (np.array_equal(a[i], ref[i], equal_nan=True) and np.array_equal(ipiv[i], piv_o[i]) and info[i] == info_o[i])]
    print(dt.__name__, cfg, "bad matrices:", len(bad), bad[:20])
    for i in bad[:3]:
        d = np.argwhere(~((a[i] == ref[i]) | (np.isnan(a[i]) & np.isnan(ref[i]))))
        print("  mat", i, "piv equal", np.array_equal(ipiv[i], piv_o[i]), "info", info[i], info_o[i], "ndiff", len(d), "first", d[:5].tolist())
        for r, c in d[:5]:
            print("    ", r, c, repr(a[i][r, c]), repr(ref[i][r, c]), a[i][r, c].view(np.uint32 if dt == np.float32 else np.uint64), ref[i][r, c].view(np.uint32 if dt == np.float32 else np.uint64))

for dt in (np.float32, np.float64):
    for cfg in (int(c) for c in os.environ.get('DBG_CFGS', '128,129').split(',')):
        run(dt, cfg)
