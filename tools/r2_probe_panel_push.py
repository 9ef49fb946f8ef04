"""Fourth-generation panel kernel (panel_push.cu, option panel_cluster = 3) against the third
(panel_blocked.cu, panel_cluster = 2): pivots and L\\U must be bit-identical on every shape,
including ragged widths, ties, zero columns and NaN; then timings and the per-phase cycles.

    python tools/r2_probe_panel_push.py [check] [time] [phases] [getrf]
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lair_b200 import _ffi  # noqa: E402

L = _ffi.lib()


def stream():
    return torch.cuda.current_stream().cuda_stream


def out(**kw):
    print(json.dumps(kw), flush=True)


def run_panel(pfx, a0, gen, rpt=0, w64=1):
    _ffi.set_option("panel_cluster", gen)
    _ffi.set_option("panel_rpt", rpt)
    _ffi.set_option("panel_w64", w64)
    m, w = a0.shape
    a = a0.clone()
    k = min(m, w)
    ipiv = torch.full((k,), -7, dtype=torch.int32, device="cuda")
    info = torch.full((1,), -7, dtype=torch.int32, device="cuda")
    fn = getattr(L, f"lair_b200_{pfx}getrf_dev")
    _ffi.check(fn(m, w, a.data_ptr(), w, ipiv.data_ptr(), info.data_ptr(), stream()))
    torch.cuda.synchronize()
    return a, ipiv, int(info.item())


def same_bits(x, y):
    it = torch.int64 if x.dtype == torch.float64 else torch.int32
    return bool(torch.equal(x.view(it), y.view(it)))


def sec_check():
    bad = 0
    cases = 0
    for dt, pfx in ((torch.float64, "d"), (torch.float32, "s")):
        g = torch.Generator(device="cuda")
        g.manual_seed(11)
        rows = [64, 200, 512, 513, 1000, 1024, 2048, 4096, 5000, 8192] + ([12000, 16384] if pfx == "s" else [])
        for m in rows:
            for w in (1, 7, 8, 9, 16, 20, 31, 32, 33, 40, 48, 63, 64):
                if w > m:
                    continue
                kinds = ["rand", "ties", "zerocol", "nan", "intvals"] if m in (200, 513, 4096, 8192) else ["rand"]
                for kind in kinds:
                    a0 = torch.rand(m, w, dtype=dt, device="cuda", generator=g) * 10
                    if kind == "ties":  # many equal magnitudes: the lowest row must win
                        a0 = torch.randint(-3, 4, (m, w), device="cuda", generator=g).to(dt)
                    elif kind == "zerocol":
                        a0[:, w // 2] = 0
                        if w > 3:
                            a0[:, 1] = a0[:, 0] * 2  # dependent column: exact zero pivot only sometimes; still must agree
                    elif kind == "nan":
                        a0[m // 3, w // 3] = float("nan")
                        a0[m - 1, 0] = float("inf")
                    elif kind == "intvals":
                        a0 = torch.randint(-100, 101, (m, w), device="cuda", generator=g).to(dt)
                    ref, piv_r, info_r = run_panel(pfx, a0, 2)
                    for rpt in (2, 1, 4):
                        cases += 1
                        got, piv_g, info_g = run_panel(pfx, a0, 3, rpt=rpt)
                        ok = torch.equal(piv_r, piv_g) and info_r == info_g and same_bits(ref, got)
                        if not ok:
                            bad += 1
                            nd = int((ref.view(torch.int64 if dt == torch.float64 else torch.int32) != got.view(torch.int64 if dt == torch.float64 else torch.int32)).sum())
                            pd = (piv_r != piv_g).nonzero().flatten()[:4].tolist()
                            out(check="MISMATCH", dtype=pfx, m=m, w=w, kind=kind, rpt=rpt, info=(info_r, info_g), piv_diff_at=pd, lu_words_diff=nd)
                            if bad > 20:
                                out(check="too many mismatches, stopping")
                                return
    out(check="panel_push vs panel_blocked", cases=cases, mismatches=bad)
    _ffi.set_option("panel_cluster", 3)
    _ffi.set_option("panel_rpt", 0)


def timeit(fn, setup, reps=7, warm=2):
    for _ in range(warm):
        setup()
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        setup()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))


def sec_time():
    for dt, pfx in ((torch.float64, "d"), (torch.float32, "s")):
        fn = getattr(L, f"lair_b200_{pfx}getrf_dev")
        for m in (512, 1024, 2048, 4096, 8192, 16384):
            for w in (32, 64):
                a0 = torch.rand(m, w, dtype=dt, device="cuda")
                a = a0.clone()
                ipiv = torch.empty(w, dtype=torch.int32, device="cuda")
                info = torch.empty(1, dtype=torch.int32, device="cuda")
                for gen, rpt in ((2, 2), (3, 1), (3, 2), (3, 4)):
                    _ffi.set_option("panel_cluster", gen)
                    _ffi.set_option("panel_rpt", rpt)
                    best, med = timeit(lambda: _ffi.check(fn(m, w, a.data_ptr(), w, ipiv.data_ptr(), info.data_ptr(), stream())),
                                       setup=lambda: a.copy_(a0))
                    out(bench=f"{pfx}panel", gen=gen, rpt=rpt, m=m, w=w, us_best=best * 1e3, us_med=med * 1e3, us_per_column=best * 1e3 / w)
    _ffi.set_option("panel_cluster", 3)
    _ffi.set_option("panel_rpt", 0)


def sec_phases():
    buf = (ctypes.c_longlong * 8)()
    names = ["candidate", "syncthreads", "cta_cand_push", "wait", "decide", "update"]
    for dt, pfx in ((torch.float64, "d"), (torch.float32, "s")):
        fn = getattr(L, f"lair_b200_{pfx}getrf_dev")
        for gen in (2, 3):
            for m, w in ((1024, 32), (8192, 32), (4096, 64)):
                _ffi.set_option("panel_cluster", gen)
                a0 = torch.rand(m, w, dtype=dt, device="cuda")
                a = a0.clone()
                ipiv = torch.empty(w, dtype=torch.int32, device="cuda")
                info = torch.empty(1, dtype=torch.int32, device="cuda")
                _ffi.set_option("panel_timing", 1)
                _ffi.check(fn(m, w, a.data_ptr(), w, ipiv.data_ptr(), info.data_ptr(), stream()))
                torch.cuda.synchronize()
                _ffi.check(L.lair_b200_debug_panel_timing(buf, 1))
                for _ in range(3):
                    a.copy_(a0)
                    _ffi.check(fn(m, w, a.data_ptr(), w, ipiv.data_ptr(), info.data_ptr(), stream()))
                torch.cuda.synchronize()
                _ffi.check(L.lair_b200_debug_panel_timing(buf, 1))
                _ffi.set_option("panel_timing", 0)
                cols = max(1, buf[6])
                out(bench=f"{pfx}panel_phases", gen=gen, m=m, w=w, columns=int(buf[6]),
                    cycles_per_column={n: round(buf[i] / cols, 1) for i, n in enumerate(names)},
                    chain_cycles_per_column=round(sum(buf[i] for i in range(6)) / cols, 1), kernel_cycles_per_launch=buf[7] / 3)
    _ffi.set_option("panel_cluster", 3)


def sec_coarse():
    """panel_timing = 2: where the launch spends its cycles outside the column loop (fourth generation only)."""
    buf = (ctypes.c_longlong * 8)()
    names = ["stage_in", "window_load", "column_loops", "last_rest_wait", "u12_solve", "rank8_update"]
    _ffi.set_option("panel_cluster", 3)
    for dt, pfx in ((torch.float64, "d"), (torch.float32, "s")):
        fn = getattr(L, f"lair_b200_{pfx}getrf_dev")
        for rpt in (2, 1):
            for m, w in ((1024, 32), (8192, 32), (4096, 64)):
                _ffi.set_option("panel_rpt", rpt)
                a0 = torch.rand(m, w, dtype=dt, device="cuda")
                a = a0.clone()
                ipiv = torch.empty(w, dtype=torch.int32, device="cuda")
                info = torch.empty(1, dtype=torch.int32, device="cuda")
                _ffi.set_option("panel_timing", 2)
                _ffi.check(fn(m, w, a.data_ptr(), w, ipiv.data_ptr(), info.data_ptr(), stream()))
                torch.cuda.synchronize()
                _ffi.check(L.lair_b200_debug_panel_timing(buf, 1))
                for _ in range(3):
                    a.copy_(a0)
                    _ffi.check(fn(m, w, a.data_ptr(), w, ipiv.data_ptr(), info.data_ptr(), stream()))
                torch.cuda.synchronize()
                _ffi.check(L.lair_b200_debug_panel_timing(buf, 1))
                _ffi.set_option("panel_timing", 0)
                per = {n: round(buf[i] / 3) for i, n in enumerate(names)}
                per["write_out_and_exit"] = round(buf[7] / 3) - sum(per.values())
                out(bench=f"{pfx}panel_coarse", rpt=rpt, m=m, w=w, cycles_per_launch=per, kernel_cycles_per_launch=round(buf[7] / 3))
    _ffi.set_option("panel_rpt", 0)


def sec_fine():
    """panel_timing = 3: inside warp 0's exchange phase."""
    buf = (ctypes.c_longlong * 8)()
    names = ["lds_argmax", "expect_tx", "window_push", "bar_wait_recip", "header_push", "mbar_wait"]
    _ffi.set_option("panel_cluster", 3)
    for dt, pfx in ((torch.float64, "d"), (torch.float32, "s")):
        fn = getattr(L, f"lair_b200_{pfx}getrf_dev")
        for m, w in ((1024, 32), (8192, 32)):
            a0 = torch.rand(m, w, dtype=dt, device="cuda")
            a = a0.clone()
            ipiv = torch.empty(w, dtype=torch.int32, device="cuda")
            info = torch.empty(1, dtype=torch.int32, device="cuda")
            _ffi.set_option("panel_timing", 3)
            _ffi.check(fn(m, w, a.data_ptr(), w, ipiv.data_ptr(), info.data_ptr(), stream()))
            torch.cuda.synchronize()
            _ffi.check(L.lair_b200_debug_panel_timing(buf, 1))
            for _ in range(3):
                a.copy_(a0)
                _ffi.check(fn(m, w, a.data_ptr(), w, ipiv.data_ptr(), info.data_ptr(), stream()))
            torch.cuda.synchronize()
            _ffi.check(L.lair_b200_debug_panel_timing(buf, 1))
            _ffi.set_option("panel_timing", 0)
            cols = max(1, buf[6])
            out(bench=f"{pfx}panel_fine", m=m, w=w, columns=int(buf[6]), cycles_per_column={n: round(buf[i] / cols, 1) for i, n in enumerate(names)})


def sec_getrf():
    for dt, pfx in ((torch.float64, "d"), (torch.float32, "s")):
        fn = getattr(L, f"lair_b200_{pfx}getrf_dev")
        for n in (2048, 4096, 8192, 16384):
            a0 = torch.rand(n, n, dtype=dt, device="cuda")
            a = a0.clone()
            ipiv = torch.empty(n, dtype=torch.int32, device="cuda")
            info = torch.empty(1, dtype=torch.int32, device="cuda")
            res = {}
            for gen in (2, 3):
                _ffi.set_option("panel_cluster", gen)
                best, med = timeit(lambda: _ffi.check(fn(n, n, a.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), stream())),
                                   setup=lambda: a.copy_(a0), reps=5)
                res[gen] = (a.clone(), ipiv.clone())
                out(bench=f"{pfx}getrf", gen=gen, n=n, ms_best=best, ms_med=med, tflops=2 / 3 * n ** 3 / best * 1e-9)
            out(bench=f"{pfx}getrf_same", n=n, pivots_equal=bool(torch.equal(res[2][1], res[3][1])), lu_bits_equal=same_bits(res[2][0], res[3][0]))
            del res
    _ffi.set_option("panel_cluster", 3)


if __name__ == "__main__":
    out(device=torch.cuda.get_device_name(0))
    secs = sys.argv[1:] or ["check", "time", "phases", "getrf"]
    for s in secs:
        globals()["sec_" + s]()
