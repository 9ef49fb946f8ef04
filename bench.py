#!/usr/bin/env python
"""bench.py -- the hot path's headline benchmark (driver contract, see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c4|c2|c3]

Metric (BASELINE.json): getrf f64 GFLOP/s with the 2/3 n^3 convention + batched LU matrices/s.
ONE headline workload at every N (so the driver's 1 -> 8 curve divides like by like): BASELINE configs[3]
("c4"), a single f64 LU of order n = 65 536 -- on one GPU through lair_b200_dgetrf_dev (34.4 GB + a pristine
copy fit one 180 GB B200), on N > 1 GPUs through the 1-D block-cyclic NCCL path (lair_b200_dgetrf_mg_dev).
`scaling` is "strong" everywhere.  One step = restore the matrix from the pristine copy + factor it.

* `value`  : device-resident (inputs already in HBM), CUDA events, max over ranks.
* `e2e`    : the same factorization through the public host API with pinned HOST buffers: H2D of A, getrf, D2H of
             L\\U and the pivots inside the timed region (N = 1: lair_b200.lapack.getrf -> lair_b200_dgetrf;
             N > 1: each rank's column slab up, lair_b200.multigpu.getrf_mg, the slab of L\\U back down).
* `roofline`: the dominant kernel (DMMA GEMM trailing update): algorithmic flops / its device time measured live
             with CUDA events on the launching stream (lair_b200_profile_*), against the measured 37.0 TFLOP/s.
* `cpu_baseline`: the oracle (C++ restatement of the reference's single-threaded algorithm) on a bounded sample,
             rank 0, N = 1 only.
* extra keys in the SAME line -- the other BASELINE configs, each with its own roofline / e2e:
     N = 1: `c2` (getrf + getrs f64 n = 8192, 64 RHS; getrf e2e with pinned AND pageable buffers),
            `c3_f64`, `c3_f32` (10^6 x 32x32 batched LU: mats/s, GB/s, fraction of the measured HBM copy bandwidth),
            `c5a` (262 144 x 1024 f32), `c5b` (16 384^2 f64) with the device-side backward error;
     N > 1: `c3_f64`, `c3_f32` with the batch sharded over the ranks (no collective).
* `--impl reference`: times the reference's CPU algorithm (oracle port; the reference is Rust and cannot be built
             in this image) on the host cores; same JSON shape.
`--workload c2` / `--workload c3` print those workloads as stand-alone lines (ncu captures, tuning).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_C2, NRHS_C2 = 8192, 64
FP64_PEAK_TFLOPS = 37.0   # measured on this pool's B200: DMMA m8n8k4 issue-rate microbench,
                          # profiles/r1_microbench_fp64_peak.jsonl (= 148 SMs x 64 FMA/clk x 1.965 GHz);
                          # cuBLAS DGEMM 8192^3 reaches 35.5 on the same box (profiles/r1_probe_first_contact.jsonl)
FP64_PEAK_SOURCE = ("FP64 DMMA issue-rate microbench on this pool's B200 (profiles/r1_microbench_fp64_peak.jsonl); "
                    "MEASURED_PEAKS.json has no FP64 entry; cuBLAS DGEMM reaches 35.5; nominal 40 is reported beside it")


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.proc = None
        self.lines = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": float(max(power))}


def _dist_setup(n_gpus: int, force: bool = False):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 or force:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        os.environ.setdefault("RANK", "0")
        os.environ.setdefault("WORLD_SIZE", "1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(local)
    return rank, world, local


def _max_over_ranks(ms: float, world: int) -> float:
    if world == 1:
        return ms
    import torch
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _barrier(world: int):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


# ------------------------------------------------------------------------------------------------
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the batched kernel over 10^6 matrices,
# from the committed `ncu --set full` capture (profiles/r2f_ncu_full_metrics.txt): f32 4.277 + 4.170 GB
# against 8.32 GB algorithmic, f64 9.050 + 8.266 GB against 16.51 GB.
NCU_TRAFFIC_C3 = {"f32": 4277404000 + 4169983000, "f64": 9049625000 + 8265569000}
NCU_TRAFFIC_C3_SOURCE = "from capture profiles/r2f_ncu_full_metrics.txt (ncu --set full of the batched kernel at this size; not measured by this run)"
BATCHED_KERNEL_NAME = "batched_lu32 (batched_lu5.cu / batched_lu4.cu, chosen by batched_cfg)"
# same counters for the largest DMMA GEMM launch of the c2 step (7808 x 7680 x 128; profiles/r2f_ncu_full_metrics.txt):
# 545.9 MB read + 459.2 MB written against 959 MB of algorithmic C read + write + 16 MB of A and B panels.
NCU_TRAFFIC_C2_GEMM = 545877760 + 459216128
# and for the first (largest) trailing update of n = 65 536 (65280 x 65280 x 256 in place; profiles/r2f_ncu_dgemm_n65536_summary.txt):
# 51.36 GB read + 34.06 GB written against 68.45 GB algorithmic (C read + write 68.18 GB, A and B panels 0.27 GB): the 134 MB
# A panel does not stay in the 126 MB L2 under the C stream and is read again for each of the 127 strips of 8 tile columns.
NCU_TRAFFIC_C4_GEMM = 51357351000 + 34058381000


def cpu_sample_c2(n_sample: int = 4096, nrhs: int = NRHS_C2, seed: int = 1):
    """The oracle on a bounded sample of the dense workload: n_sample x n_sample f64 getrf + nrhs
    single-RHS getrs calls (the reference has no multi-RHS getrs), ONE core."""
    import oracle
    rng = np.random.default_rng(seed)
    a = rng.uniform(0, 10, size=(n_sample, n_sample))
    b = rng.uniform(0, 10, size=(n_sample, nrhs))
    tg, ts, _, _ = oracle.time_dgetrf_dgetrs(a, b)
    flops = 2.0 / 3.0 * n_sample ** 3 + 2.0 * n_sample ** 2 * nrhs
    return flops / (tg + ts) * 1e-9, tg, ts, f"n={n_sample} f64 getrf + {nrhs} single-RHS getrs (oracle port of lair, -O2 -ffp-contract=off), 1 thread"


def run_reference(args, rank: int, world: int):
    """--impl reference: the reference's own CPU algorithm on the host cores (oracle port --
    lair is single-threaded, so 1 core is all it can use)."""
    if rank != 0:
        return
    for _ in range(min(args.warmup, 1)):
        cpu_sample_c2(512)
    # each step is a bounded sample; with many steps the sample shrinks so the run ends within minutes
    ref_n = args.ref_n if args.steps <= 8 else min(args.ref_n, 3072) if args.steps <= 24 else min(args.ref_n, 2048)
    vals = []
    t0 = time.perf_counter()
    sample = ""
    for _ in range(args.steps):
        g, tg, ts, sample = cpu_sample_c2(ref_n)
        vals.append((g, tg + ts))
    wall = time.perf_counter() - t0
    gflops = float(np.mean([v[0] for v in vals]))
    line = {
        "impl": "reference", "metric": "getrf_f64_gflops", "value": gflops, "unit": "GFLOP/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic uniform[0,10)",
        "config": {"workload": f"c4: single getrf f64 n={args.n} (bounded sample: n={ref_n} getrf + {NRHS_C2} getrs on ONE host core; "
                               "the reference is O(n^3) single-threaded scalar code, n=65536 would take ~35 h)", "inputs": "host"},
        "cpu_baseline": {"value": gflops, "unit": "GFLOP/s", "cores": 1, "kind": "port", "sample": sample,
                         "host_cores_available": os.cpu_count()},
        "e2e": {"value": gflops, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def _time_host_call(fn, steps: int, restore=None) -> float:
    """Seconds per call of a host-API entry (wall clock: the call returns when the result is back in host memory)."""
    t = 0.0
    for _ in range(steps):
        if restore:
            restore()
        t0 = time.perf_counter()
        fn()
        t += time.perf_counter() - t0
    return t / steps


def measure_c2(args, rank: int, world: int, local: int, sampler=None) -> dict:
    """BASELINE configs[1]: getrf + getrs, f64, n = 8192, 64 right-hand sides, one system per rank."""
    import torch
    from lair_b200 import _ffi
    import lair_b200

    L = _ffi.lib()
    n, nrhs = N_C2, NRHS_C2
    flops_step = 2.0 / 3.0 * n ** 3 + 2.0 * n * n * nrhs
    stream = torch.cuda.current_stream().cuda_stream
    steps = args.steps

    gen = torch.Generator(device="cuda")
    gen.manual_seed(1 + rank)
    ncopies = max(1, min(steps, 12))
    a0 = torch.rand(n, n, dtype=torch.float64, device="cuda", generator=gen) * 10
    b0 = torch.rand(n, nrhs, dtype=torch.float64, device="cuda", generator=gen) * 10
    a_bufs = [a0.clone() for _ in range(ncopies)]
    b_bufs = [b0.clone() for _ in range(ncopies)]
    ipiv = torch.empty(n, dtype=torch.int32, device="cuda")
    info = torch.empty(1, dtype=torch.int32, device="cuda")

    def step(i):
        a, b = a_bufs[i % ncopies], b_bufs[i % ncopies]
        _ffi.check(L.lair_b200_dgetrf_dev(n, n, a.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), stream))
        _ffi.check(L.lair_b200_dgetrs_dev(n, nrhs, a.data_ptr(), n, ipiv.data_ptr(), b.data_ptr(), nrhs, stream))

    def restore():
        for a, b in zip(a_bufs, b_bufs):
            a.copy_(a0)
            b.copy_(b0)

    for i in range(args.warmup):
        step(i)
    restore()
    _barrier(world)
    if sampler is not None:
        sampler.start()
    launches0 = _ffi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _barrier(world)
    e0.record()
    for i in range(steps):
        if i and i % ncopies == 0:
            restore()  # only when steps > 12 buffers (not in the default run)
        step(i)
    e1.record()
    _barrier(world)
    ms_total = _max_over_ranks(e0.elapsed_time(e1), world)
    _ffi.check_fault(stream)  # a timed-out device-side wait would invalidate the run
    launches = _ffi.launch_count() - launches0
    clocks = sampler.stop() if sampler is not None else None
    ms_step = ms_total / steps
    value = world * flops_step / ms_step * 1e-6  # GFLOP/s, whole job

    # correctness of what was timed: residual of the last solved system, on the device
    x = b_bufs[(steps - 1) % ncopies]
    res = float(torch.linalg.norm(a0 @ x - b0) / (torch.linalg.norm(a0) * torch.linalg.norm(x) * n * 2.0 ** -53))
    info_val = int(info.item())

    if args.ncu_step:  # one warm step inside the profiler range, for the committed ncu launch list
        restore()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step(0)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()

    # getrf-only / getrs-only split (separate pass, same stream, same conditions as the timed steps)
    restore()
    torch.cuda.synchronize()
    pe0, pe1, pe2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    pe0.record()
    _ffi.check(L.lair_b200_dgetrf_dev(n, n, a_bufs[0].data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), stream))
    pe1.record()
    _ffi.check(L.lair_b200_dgetrs_dev(n, nrhs, a_bufs[0].data_ptr(), n, ipiv.data_ptr(), b_bufs[0].data_ptr(), nrhs, stream))
    pe2.record()
    torch.cuda.synchronize()
    getrf_ms, getrs_ms = pe0.elapsed_time(pe1), pe1.elapsed_time(pe2)

    # live per-kernel-family timing: every launch bracketed by CUDA events on its own stream (this
    # pass serialises the two streams' kernels, so its family sums exceed the overlapped step time)
    restore()
    torch.cuda.synchronize()
    _ffi.profile_begin()
    _ffi.check(L.lair_b200_dgetrf_dev(n, n, a_bufs[0].data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), stream))
    _ffi.check(L.lair_b200_dgetrs_dev(n, nrhs, a_bufs[0].data_ptr(), n, ipiv.data_ptr(), b_bufs[0].data_ptr(), nrhs, stream))
    prof = _ffi.profile_end()
    gemm = prof["gemm"]
    gemm_tflops = gemm["work"] / gemm["ms"] * 1e-9 if gemm["ms"] > 0 else 0.0
    kernel_share = {k: round(v["ms"], 3) for k, v in prof.items() if v["launches"]}
    del a_bufs, b_bufs

    # end to end through the public host API
    e2e = e2e_getrf_pinned = e2e_getrf_pageable = None
    if not args.no_e2e:
        a_host = torch.empty(n, n, dtype=torch.float64).pin_memory()
        b_host = torch.empty(n, nrhs, dtype=torch.float64).pin_memory()
        a_host.copy_(a0)
        b_host.copy_(b0)
        a_np, b_np = a_host.numpy(), b_host.numpy()
        e2e_steps = max(2, min(steps, 5))
        lair_b200.equation.solve(a_np, b_np)  # warm-up (allocates the device pool)
        _barrier(world)
        box = {}
        t_e2e = _time_host_call(lambda: box.__setitem__("x", lair_b200.equation.solve(a_np, b_np)), e2e_steps)
        t_e2e = _max_over_ranks(t_e2e * 1e3, world) * 1e-3
        e2e = {"value": world * flops_step / t_e2e * 1e-9, "unit": "GFLOP/s", "ms_per_step": t_e2e * 1e3,
               "h2d_bytes_per_step": int(a_np.nbytes + b_np.nbytes), "d2h_bytes_per_step": int(box["x"].nbytes + 4),
               "api": "lair_b200.equation.solve -> lair_b200_dgesv (pinned host buffers)"}
        # the getrf contract itself (getrf.rs:12 overwrites the caller's array): H2D of A, factor, D2H of L\U + pivots
        gflops_getrf = 2.0 / 3.0 * n ** 3
        work_pin = torch.empty(n, n, dtype=torch.float64).pin_memory()
        w_np = work_pin.numpy()
        w_np[...] = a_np
        lair_b200.lapack.getrf(w_np)
        t = _time_host_call(lambda: lair_b200.lapack.getrf(w_np), e2e_steps, restore=lambda: np.copyto(w_np, a_np))
        t = _max_over_ranks(t * 1e3, world) * 1e-3
        e2e_getrf_pinned = {"value": world * gflops_getrf / t * 1e-9, "unit": "GFLOP/s", "ms_per_step": t * 1e3,
                            "h2d_bytes_per_step": int(a_np.nbytes), "d2h_bytes_per_step": int(a_np.nbytes + 8 * n + 8),
                            "api": "lair_b200.lapack.getrf -> lair_b200_dgetrf (pinned host array, factored in place)"}
        w_page = np.empty((n, n), dtype=np.float64)  # an ordinary ndarray: pageable
        w_page[...] = a_np
        lair_b200.lapack.getrf(w_page)
        t = _time_host_call(lambda: lair_b200.lapack.getrf(w_page), e2e_steps, restore=lambda: np.copyto(w_page, a_np))
        t = _max_over_ranks(t * 1e3, world) * 1e-3
        e2e_getrf_pageable = {"value": world * gflops_getrf / t * 1e-9, "unit": "GFLOP/s", "ms_per_step": t * 1e3,
                              "h2d_bytes_per_step": int(a_np.nbytes), "d2h_bytes_per_step": int(a_np.nbytes + 8 * n + 8),
                              "api": "lair_b200.lapack.getrf -> lair_b200_dgetrf (pageable numpy array, factored in place)"}
        del a_host, b_host, work_pin, w_page

    return {
        "metric": "getrf_f64_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic uniform[0,10), random seed per rank",
        "config": {"workload": f"c2: getrf+getrs f64 n={n} nrhs={nrhs} per GPU (flops = 2/3 n^3 + 2 n^2 nrhs)",
                   "l2": "inputs (512 MiB per system, a fresh buffer per step) exceed the 126 MB L2",
                   "nb": _ffi.get_option("nb") or "auto by remaining size", "sharding": "independent systems per rank, no collective"},
        "getrf_ms": getrf_ms, "getrs_ms": getrs_ms, "getrf_gflops": 2.0 / 3.0 * n ** 3 / getrf_ms * 1e-6,
        "frac_of_fp64_peak": value * 1e-3 / (FP64_PEAK_TFLOPS * world),
        "residual_scaled": res, "info": info_val,
        "roofline": {"bound": "tensor", "kernel": "dgemm_minus_kernel (DMMA m8n8k4 trailing update)",
                     "achieved": gemm_tflops, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                     "frac": gemm_tflops / FP64_PEAK_TFLOPS, "peak_source": FP64_PEAK_SOURCE,
                     "launches": gemm["launches"], "kernel_ms_in_step": gemm["ms"],
                     "traffic": NCU_TRAFFIC_C2_GEMM,
                     "traffic_unit": "bytes of the step's largest launch (M x N x K = 7808 x 7680 x 128; "
                                     "dram__bytes_read.sum + dram__bytes_write.sum), algorithmic C read + write = 959 MB",
                     "traffic_source": "from capture profiles/r2f_ncu_full_metrics.txt (ncu --set full of that launch; not measured by this run)",
                     "kernel_ms_by_family": kernel_share},
        "e2e": e2e, "e2e_getrf_pinned": e2e_getrf_pinned, "e2e_getrf_pageable": e2e_getrf_pageable,
        "gpu_launches": int(launches), "clocks": clocks,
    }


def measure_c3(args, dtype: str, rank: int, world: int, local: int, sampler=None, cpu: bool = True) -> dict:
    """BASELINE configs[2]: batched 32x32 LU, 10^6 matrices sharded over the ranks (HBM-bound path)."""
    import torch
    from lair_b200 import _ffi, sharding
    import lair_b200

    L = _ffi.lib()
    peaks, peak_src = _peaks()
    dt, pfx, bpm = (torch.float64, "d", 16512) if dtype == "f64" else (torch.float32, "s", 8320)
    total = 1_000_000
    _, batch = sharding.batch_slice(total, rank, world)  # contiguous slices of the batch, no collective
    stream = torch.cuda.current_stream().cuda_stream
    gen = torch.Generator(device="cuda")
    gen.manual_seed(3 + rank)
    a0 = torch.rand(batch, 32, 32, dtype=dt, device="cuda", generator=gen) * 10
    a = a0.clone()
    ipiv = torch.empty(batch, 32, dtype=torch.int32, device="cuda")
    info = torch.empty(batch, dtype=torch.int32, device="cuda")
    fn = getattr(L, f"lair_b200_{pfx}getrf_batched_dev")
    steps = max(3, min(args.steps, 10))

    def step():
        _ffi.check(fn(batch, 32, a.data_ptr(), ipiv.data_ptr(), info.data_ptr(), stream))

    for _ in range(max(3, args.warmup)):
        a.copy_(a0)
        step()
    if sampler is not None:
        sampler.start()
    # time each step separately so the restore copy (a <- a0) stays outside the timed region
    ms_total = 0.0
    launches0 = _ffi.launch_count()
    _barrier(world)
    for _ in range(steps):
        a.copy_(a0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        step()
        e1.record()
        torch.cuda.synchronize()
        ms_total += e0.elapsed_time(e1)
    _barrier(world)
    ms_total = _max_over_ranks(ms_total, world)
    launches = _ffi.launch_count() - launches0
    clocks = sampler.stop() if sampler is not None else None
    ms_step = ms_total / steps
    value = total / ms_step * 1e3
    gbs = batch * bpm / ms_step * 1e-6  # per GPU
    # what was timed is right: P A = L U on a sample, on the device in f64; no singular flags
    ns = min(batch, 4096)
    LU = a[:ns].double()
    Lm = torch.tril(LU, -1) + torch.eye(32, dtype=torch.float64, device="cuda")
    rec = Lm @ torch.triu(LU)
    pv = ipiv[:ns].cpu().numpy()
    perm = np.tile(np.arange(32), (ns, 1))
    rows = np.arange(ns)
    for j in range(32):
        p = pv[:, j]
        tmp = perm[rows, j].copy()
        perm[rows, j] = perm[rows, p]
        perm[rows, p] = tmp
    PA = torch.gather(a0[:ns].double(), 1, torch.from_numpy(perm).cuda()[:, :, None].expand(-1, -1, 32))
    eps = 2.0 ** -53 if dtype == "f64" else 2.0 ** -24
    be = float((torch.linalg.norm((PA - rec).reshape(ns, -1), dim=1) / (32 * eps * torch.linalg.norm(PA.reshape(ns, -1), dim=1))).max())
    n_singular = int((info >= 0).sum().item())
    e2e = None
    if not args.no_e2e:
        eb = min(batch, 200_000)
        host = torch.empty(eb, 32, 32, dtype=dt).pin_memory()
        host.copy_(a0[:eb])
        h_np = host.numpy()
        work_t = torch.empty_like(host).pin_memory()  # the call factors in place: a pinned working copy, restored untimed
        work = work_t.numpy()
        work[...] = h_np
        lair_b200.lapack.getrf_batched(work)
        box = {}
        t = _time_host_call(lambda: box.__setitem__("r", lair_b200.lapack.getrf_batched(work)), 3, restore=lambda: np.copyto(work, h_np))
        t = _max_over_ranks(t * 1e3, world) * 1e-3
        p, i_ = box["r"]  # H2D of the batch, factorization, D2H of L\U + pivots + info
        e2e = {"value": eb * world / t, "unit": "mats/s", "h2d_bytes_per_step": int(h_np.nbytes),
               "d2h_bytes_per_step": int(h_np.nbytes + p.nbytes + i_.nbytes), "sample": f"{eb} matrices per call per rank",
               "api": f"lair_b200.lapack.getrf_batched -> lair_b200_{pfx}getrf_batched (pinned host buffers)"}
        del host, work_t
    cpu_b = None
    if cpu and rank == 0 and world == 1 and not args.no_cpu:
        import oracle
        smp = a0[:min(batch, args.ref_batch)].cpu().numpy()
        t0 = time.perf_counter()
        oracle.getrf_batched(smp)
        t = time.perf_counter() - t0
        cpu_b = {"value": len(smp) / t, "unit": "mats/s", "cores": 1, "kind": "port",
                 "sample": f"{len(smp)} matrices of the same batch ({t:.1f} s), 1 thread"}
    return {
        "metric": "batched_lu32_mats_per_s", "value": value, "unit": "mats/s", "n_gpus": world, "steps": steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": dtype, "data": "synthetic uniform[0,10)",
        "config": {"workload": f"c3: batched getrf {dtype} 10^6 x 32x32, batch sharded over ranks ({batch} per rank, no collective)",
                   "l2": "per-rank input exceeds L2 at N<=4; restore copy outside the timed region",
                   "batched_cfg": _ffi.get_option("batched_cfg")},
        "mats_per_s_per_gpu": value / world, "backward_error_max_sample": be, "singular_flags": n_singular,
        "roofline": {"bound": "hbm", "kernel": BATCHED_KERNEL_NAME, "achieved": gbs, "peak": peaks["hbm_gbs"],
                     "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"], "peak_source": peak_src,
                     "traffic": NCU_TRAFFIC_C3.get(dtype) if batch == 1_000_000 else None,
                     "traffic_unit": "bytes per launch of 10^6 matrices (dram__bytes_read.sum + dram__bytes_write.sum)",
                     "traffic_source": NCU_TRAFFIC_C3_SOURCE,
                     "algorithmic_bytes_per_launch": int(batch * bpm), "algorithmic_bytes_per_matrix": bpm},
        "cpu_baseline": cpu_b, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }


def measure_c5(args, which: str) -> dict:
    """BASELINE configs[4]: tall-skinny f32 262 144 x 1024 (c5a, panel-dominated) and f64 n = 16 384 (c5b), one GPU."""
    import torch
    from lair_b200 import _ffi
    import devcheck

    L = _ffi.lib()
    m, n, dt, pfx = (262144, 1024, torch.float32, "s") if which == "c5a" else (16384, 16384, torch.float64, "d")
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5 if which == "c5a" else 6)
    a0 = torch.rand(m, n, dtype=dt, device="cuda", generator=gen) * 10
    a = torch.empty_like(a0)
    k = min(m, n)
    ipiv = torch.empty(k, dtype=torch.int32, device="cuda")
    info = torch.empty(1, dtype=torch.int32, device="cuda")
    fn = getattr(L, f"lair_b200_{pfx}getrf_dev")
    stream = torch.cuda.current_stream().cuda_stream
    steps = max(2, min(args.steps, 5))
    ts = []
    for i in range(3 + steps):
        a.copy_(a0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _ffi.check(fn(m, n, a.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), stream))
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    _ffi.check_fault(stream)
    ms = float(np.mean(ts))
    # live per-family times of one more factorization
    a.copy_(a0)
    torch.cuda.synchronize()
    _ffi.profile_begin()
    _ffi.check(fn(m, n, a.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), stream))
    prof = _ffi.profile_end()
    flops = m * n * n - n ** 3 / 3.0
    be = devcheck.backward_error_dev(a0, a, ipiv)
    esz = 4 if dt == torch.float32 else 8
    out = {"workload": f"{which}: getrf {'f32' if esz == 4 else 'f64'} {m} x {n}, one GPU (flops = m n^2 - n^3/3)", "ms": ms, "steps": steps, "warmup": 3,
           "gflops": flops / ms * 1e-6, "backward_error": be, "backward_error_bound": 0.5, "info": int(info.item()),
           "kernel_ms_by_family": {kk: round(v["ms"], 3) for kk, v in prof.items() if v["launches"]}}
    if which == "c5b":
        out["frac_of_fp64_peak"] = flops / ms * 1e-9 / FP64_PEAK_TFLOPS
        g = prof["gemm"]
        out["roofline"] = {"bound": "tensor", "kernel": "dgemm_minus_kernel", "achieved": g["work"] / g["ms"] * 1e-9 if g["ms"] > 0 else 0.0,
                           "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": (g["work"] / g["ms"] * 1e-9 / FP64_PEAK_TFLOPS) if g["ms"] > 0 else 0.0, "traffic": None}
    else:
        peaks, peak_src = _peaks()
        pn = prof["panel"]
        # panel kernels: every column block of the panel is read and written once per launch (algorithmic 2 * rows * w * sizeof)
        out["roofline"] = {"bound": "hbm", "kernel": "panel kernels (panel_blocked.cu / panel.cu)", "achieved": pn["work"] / pn["ms"] * 1e-6 if pn["ms"] > 0 else 0.0,
                           "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": (pn["work"] / pn["ms"] * 1e-6 / peaks["hbm_gbs"]) if pn["ms"] > 0 else 0.0,
                           "peak_source": peak_src, "traffic": None,
                           "note": "latency-bound (one dependent arg-max per column), reported against HBM as SURVEY 8(d) asks"}
    del a0, a
    torch.cuda.empty_cache()
    return out


def _host_mem_available() -> int:
    """Free host memory this process may pin: the smaller of the machine's available memory and the
    cgroup's remaining allowance (a container limit that psutil does not see)."""
    try:
        import psutil
        avail = int(psutil.virtual_memory().available)
    except Exception:
        return 0
    for lim_f, use_f in (("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory.current"),
                         ("/sys/fs/cgroup/memory/memory.limit_in_bytes", "/sys/fs/cgroup/memory/memory.usage_in_bytes")):
        try:
            lim = open(lim_f).read().strip()
            if lim != "max":
                avail = min(avail, int(lim) - int(open(use_f).read().strip()))
        except Exception:
            pass
    return max(avail, 0)


def run_c4_single(args, rank: int, world: int, local: int) -> dict:
    """The headline workload on ONE GPU: a single f64 LU of order n (65 536: 34.4 GB + a pristine copy) through
    lair_b200_dgetrf_dev -- the N = 1 point of the strong-scaling curve."""
    import torch
    from lair_b200 import _ffi
    import lair_b200

    L = _ffi.lib()
    n = args.n
    stream = torch.cuda.current_stream().cuda_stream
    gen = torch.Generator(device="cuda")
    gen.manual_seed(4)
    a0 = torch.empty(n, n, dtype=torch.float64, device="cuda")
    for r0 in range(0, n, 8192):  # generated by row blocks: torch.rand(...) * 10 would need a second n x n temporary
        r1 = min(n, r0 + 8192)
        a0[r0:r1] = torch.rand(r1 - r0, n, dtype=torch.float64, device="cuda", generator=gen) * 10
    a = torch.empty_like(a0)
    ipiv = torch.empty(n, dtype=torch.int32, device="cuda")
    info = torch.empty(1, dtype=torch.int32, device="cuda")
    flops = 2.0 / 3.0 * n ** 3

    def step():
        a.copy_(a0)  # in-place factorization: restore from the pristine copy (inside the timed region, ~0.3 % of a step)
        _ffi.check(L.lair_b200_dgetrf_dev(n, n, a.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), stream))

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    torch.cuda.synchronize()
    sampler.start()
    launches0 = _ffi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms_step = e0.elapsed_time(e1) / args.steps
    _ffi.check_fault(stream)
    launches = _ffi.launch_count() - launches0
    clocks = sampler.stop()
    value = flops / ms_step * 1e-6

    # size-independent parity property: || (P A - L U) x || / (||A|| ||x|| n eps) for a random x
    torch.manual_seed(99)
    x = torch.rand(n, dtype=torch.float64, device="cuda")
    y = torch.zeros(n, dtype=torch.float64, device="cuda")
    z = torch.zeros(n, dtype=torch.float64, device="cuda")
    w = a0 @ x
    bs = 4096
    for c0 in range(0, n, bs):  # y = U x
        c1 = min(n, c0 + bs)
        y[:c0] += a[:c0, c0:c1] @ x[c0:c1]
        y[c0:c1] += torch.triu(a[c0:c1, c0:c1]) @ x[c0:c1]
    for c0 in range(0, n, bs):  # z = L y
        c1 = min(n, c0 + bs)
        z[c1:] += a[c1:, c0:c1] @ y[c0:c1]
        z[c0:c1] += (torch.tril(a[c0:c1, c0:c1], -1) + torch.eye(c1 - c0, dtype=torch.float64, device="cuda")) @ y[c0:c1]
    import devcheck
    perm = torch.from_numpy(devcheck.perm_from_pivots(ipiv.cpu().numpy(), n)).cuda()
    resid = float(torch.linalg.norm(w[perm] - z) / (torch.linalg.norm(a0) * torch.linalg.norm(x) * n * 2.0 ** -53))
    info_val = int(info.item())

    # live per-kernel-family timing of one more factorization (serialises the streams)
    a.copy_(a0)
    torch.cuda.synchronize()
    _ffi.profile_begin()
    _ffi.check(L.lair_b200_dgetrf_dev(n, n, a.data_ptr(), n, ipiv.data_ptr(), info.data_ptr(), stream))
    prof = _ffi.profile_end()
    gemm = prof["gemm"]
    gemm_tflops = gemm["work"] / gemm["ms"] * 1e-9 if gemm["ms"] > 0 else 0.0
    kernel_share = {k: round(v["ms"], 3) for k, v in prof.items() if v["launches"]}
    del a
    torch.cuda.empty_cache()

    # end to end: the getrf contract through the host API -- H2D of A, factor, D2H of L\U and the pivots
    e2e = None
    if not args.no_e2e:
        nbytes = n * n * 8
        if nbytes * 1.2 < _host_mem_available():
            host = torch.empty(n, n, dtype=torch.float64).pin_memory()
            h_np = host.numpy()
            e2e_steps = 2

            def restore():
                host.copy_(a0)  # untimed: the pristine matrix back into the pinned host array
                torch.cuda.synchronize()

            restore()
            lair_b200.lapack.getrf(h_np)  # warm-up: sizes the library's device pool
            t = _time_host_call(lambda: lair_b200.lapack.getrf(h_np), e2e_steps, restore=restore)
            e2e = {"value": flops / t * 1e-9, "unit": "GFLOP/s", "ms_per_step": t * 1e3, "steps": e2e_steps,
                   "h2d_bytes_per_step": int(nbytes), "d2h_bytes_per_step": int(nbytes + 8 * n + 8),
                   "api": "lair_b200.lapack.getrf -> lair_b200_dgetrf (pinned host array factored in place: H2D of A, getrf, D2H of L\\U + pivots)"}
            del host
        else:
            e2e = {"value": None, "unit": "GFLOP/s", "skipped": f"{nbytes / 2**30:.0f} GiB pinned host array vs {_host_mem_available() / 2**30:.0f} GiB free",
                   "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    del a0
    torch.cuda.empty_cache()

    return {
        "metric": "getrf_f64_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic uniform[0,10), generated on device",
        "config": {"workload": f"c4: single getrf f64 n={n} on 1 GPU (lair_b200_dgetrf_dev; the same matrix order the N>1 lines distribute)",
                   "l2": f"matrix {n * n * 8 / 2**30:.1f} GiB exceeds L2; restored from a pristine copy inside the timed region",
                   "flops": "2/3 n^3", "nb": _ffi.get_option("nb") or "auto by remaining size"},
        "frac_of_aggregate_fp64_peak": value * 1e-3 / FP64_PEAK_TFLOPS, "aggregate_fp64_peak_tflops": FP64_PEAK_TFLOPS,
        "frac_of_nominal_40": value * 1e-3 / 40.0,
        "residual_scaled_PA_minus_LU_times_x": resid, "info": info_val,
        "roofline": {"bound": "tensor", "kernel": "dgemm_minus_kernel (DMMA m8n8k4 trailing update)", "achieved": gemm_tflops,
                     "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": gemm_tflops / FP64_PEAK_TFLOPS, "peak_source": FP64_PEAK_SOURCE,
                     "launches": gemm["launches"], "kernel_ms_in_step": gemm["ms"],
                     "traffic": NCU_TRAFFIC_C4_GEMM if n == 65536 else None,
                     "traffic_unit": "bytes of the step's largest launch (M x N x K = 65280 x 65280 x 256; dram__bytes_read.sum + "
                                     "dram__bytes_write.sum), algorithmic 68.45 GB (C read + write + the A and B panels once)",
                     "traffic_source": "from capture profiles/r2f_ncu_dgemm_n65536_summary.txt (ncu --set full of that launch: 67.7 ms, "
                                       "32.2 TFLOP/s, DMMA pipe 86.7 %; not measured by this run)",
                     "kernel_ms_by_family": kernel_share},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }


def run_c4(args, rank: int, world: int, local: int) -> dict:
    """One large f64 LU, 1-D block-cyclic columns over the ranks, NCCL panel broadcast + lookahead
    (BASELINE configs[3]; n = 65536 unless --order).  Strong scaling: the matrix is fixed, ranks split it."""
    import torch
    import torch.distributed as dist
    from lair_b200 import _ffi, multigpu, sharding

    L = _ffi.lib()
    multigpu.init()
    n, nb = args.n, args.nb
    lcols = sharding.local_cols(n, nb, rank, world)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(4 + rank)
    a0 = torch.rand(n, lcols, dtype=torch.float64, device="cuda", generator=gen) * 10
    a = torch.empty_like(a0)
    flops = 2.0 / 3.0 * n ** 3

    def step():
        a.copy_(a0)  # in-place factorization: restore the slab (6 ms of 1 s at n = 65536; inside the timed region)
        return multigpu.getrf_mg(a, n, nb)

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    _barrier(world)
    if rank == 0:
        sampler.start()
    launches0 = _ffi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _barrier(world)
    e0.record()
    for _ in range(args.steps):
        ipiv, info = step()
    e1.record()
    _barrier(world)
    ms_total = _max_over_ranks(e0.elapsed_time(e1), world)
    _ffi.check_fault(torch.cuda.current_stream().cuda_stream)  # a timed-out device-side wait would invalidate the run
    launches = _ffi.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = flops / ms_step * 1e-6  # GFLOP/s of the whole job

    # size-independent parity property: || (P A - L U) x || / (||A|| ||x|| n eps) for a random x
    torch.manual_seed(99)
    x = torch.rand(n, dtype=torch.float64, device="cuda")
    y = torch.zeros(n, dtype=torch.float64, device="cuda")
    w = torch.zeros(n, dtype=torch.float64, device="cuda")
    blocks = sharding.local_blocks(n, nb, rank, world)
    for lb, g in enumerate(blocks):
        c0, wd = g * nb, min(nb, n - g * nb)
        sub = a[:, lb * nb: lb * nb + wd]
        xs = x[c0:c0 + wd]
        y[:c0] += sub[:c0] @ xs
        y[c0:c0 + wd] += torch.triu(sub[c0:c0 + wd]) @ xs
        w += a0[:, lb * nb: lb * nb + wd] @ xs
    dist.all_reduce(y)
    dist.all_reduce(w)
    z = torch.zeros(n, dtype=torch.float64, device="cuda")
    for lb, g in enumerate(blocks):
        c0, wd = g * nb, min(nb, n - g * nb)
        sub = a[:, lb * nb: lb * nb + wd]
        ys = y[c0:c0 + wd]
        z[c0 + wd:] += sub[c0 + wd:] @ ys
        z[c0:c0 + wd] += (torch.tril(sub[c0:c0 + wd], -1) + torch.eye(wd, dtype=torch.float64, device="cuda")) @ ys
    dist.all_reduce(z)
    anorm2 = (a0 * a0).sum()
    dist.all_reduce(anorm2)
    pw = w.cpu().numpy()
    for i, p in enumerate(ipiv.cpu().numpy()):
        if i != p:
            pw[i], pw[p] = pw[p], pw[i]
    resid = float(np.linalg.norm(pw - z.cpu().numpy()) / (float(anorm2.sqrt()) * float(torch.linalg.norm(x)) * n * 2.0 ** -53))

    # end to end through the public API (lair_b200.multigpu.getrf_mg): every step uploads the rank's column slab
    # from pinned host memory, factors, and brings the slab of L\U, the pivots and info back to the host.
    # Skipped (null, with the reason) when the slabs of all ranks would not comfortably fit the host's free memory.
    e2e = None
    if not args.no_e2e:
        slab_bytes = a0.numel() * 8
        avail = _host_mem_available()
        fits = torch.tensor([1 if slab_bytes * world * 2.4 < avail else 0], device="cuda")  # an input and an output slab per rank
        dist.all_reduce(fits, op=dist.ReduceOp.MIN)
        if int(fits.item()) == 1:
            host = torch.empty(a0.shape, dtype=torch.float64).pin_memory()
            host_out = torch.empty(a0.shape, dtype=torch.float64).pin_memory()
            host.copy_(a0)
            e2e_steps = 2

            def e2e_step():
                a.copy_(host, non_blocking=True)
                piv, inf = multigpu.getrf_mg(a, n, nb)
                host_out.copy_(a, non_blocking=True)
                piv_h, inf_h = piv.cpu(), int(inf.cpu().item())  # synchronises: the slab copy above is on the same stream
                return piv_h, inf_h

            e2e_step()
            _barrier(world)
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                piv_h, _inf = e2e_step()
            _barrier(world)
            t_e2e = _max_over_ranks((time.perf_counter() - t0) / e2e_steps * 1e3, world) * 1e-3
            e2e = {"value": flops / t_e2e * 1e-9, "unit": "GFLOP/s", "ms_per_step": t_e2e * 1e3, "steps": e2e_steps,
                   "h2d_bytes_per_step": int(slab_bytes) * world, "d2h_bytes_per_step": int(slab_bytes + piv_h.numel() * 4 + 4) * world,
                   "api": "lair_b200.multigpu.getrf_mg (per-rank column slab from pinned host memory; the slab of L\\U + pivots + info read back)"}
            del host, host_out
        else:
            e2e = {"value": None, "unit": "GFLOP/s", "skipped": f"{world} x {slab_bytes / 2**30:.1f} GiB of pinned slabs "
                   f"vs {avail / 2**30:.0f} GiB of free host memory", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    del a0, a
    torch.cuda.empty_cache()
    multigpu.finalize()

    peak = FP64_PEAK_TFLOPS * world
    return {
        "metric": "getrf_f64_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic uniform[0,10), generated per rank on device",
        "config": {"workload": f"c4: single getrf f64 n={n}, 1-D block-cyclic columns nb={nb} over {world} GPUs, "
                               "NCCL panel+pivot broadcast with one block of lookahead",
                   "l2": f"per-rank slab {n * lcols * 8 / 2**30:.1f} GiB exceeds L2; slab restored from a pristine copy inside the timed region",
                   "flops": "2/3 n^3"},
        "frac_of_aggregate_fp64_peak": value * 1e-3 / peak, "aggregate_fp64_peak_tflops": peak,
        "frac_of_nominal_40": value * 1e-3 / (40.0 * world),
        "residual_scaled_PA_minus_LU_times_x": resid, "info": int(info.item()),
        "roofline": {"bound": "tensor", "kernel": "dgemm_minus_kernel (DMMA m8n8k4 trailing update)", "achieved": value * 1e-3 / world,
                     "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": value * 1e-3 / peak, "peak_source": FP64_PEAK_SOURCE,
                     "note": "whole-factorization FLOP/s per GPU (the kernel-only figure is reported by the N=1 line)", "traffic": None},
        "cpu_baseline": None,
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="lair_b200", choices=["lair_b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c2", "c3", "c4"],
                    help="default c4: one n=65536 LU (single GPU at N=1, distributed over the ranks at N>1) + the other configs as extra keys")
    ap.add_argument("--order", dest="n", type=int, default=65536, help="matrix order n of the c4 workload")
    ap.add_argument("--block", dest="nb", type=int, default=256, help="block-cyclic block width of the c4 workload")
    ap.add_argument("--dtype", default="f64", choices=["f32", "f64"])
    ap.add_argument("--ref-batch", type=int, default=100_000, help="matrices in the CPU (oracle) leg of the c3 workload")
    ap.add_argument("--ref-n", type=int, default=4096, help="sample size of the CPU (oracle) leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="c4: skip the c2 / c3 / c5 extra keys")
    ap.add_argument("--ncu-step", action="store_true",
                    help="c2: after the timed region run ONE more step between cudaProfilerStart/Stop "
                         "(for `ncu --profile-from-start off`; numbers printed by such a run are not bench values)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl != "reference":
        args.warmup = 3  # timing rule: at least 3 warm-up steps

    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        run_reference(args, rank, world)
        return
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    rank, world, local = _dist_setup(args.gpus, force=(args.workload == "c4" and world_env > 1))
    from lair_b200 import _ffi
    _ffi.check(_ffi.lib().lair_b200_init(local))
    try:
        if args.workload == "c2":
            line = measure_c2(args, rank, world, local, ClockSampler(local) if rank == 0 else None)
            if rank == 0 and world == 1 and not args.no_cpu:
                g, tg, ts, sample = cpu_sample_c2(args.ref_n)
                line["cpu_baseline"] = {"value": g, "unit": "GFLOP/s", "cores": 1, "kind": "port", "sample": sample, "host_cores_available": os.cpu_count()}
        elif args.workload == "c3":
            line = measure_c3(args, args.dtype, rank, world, local, ClockSampler(local) if rank == 0 else None)
        elif world == 1:
            line = run_c4_single(args, rank, world, local)
            if not args.no_extras:
                line["c2"] = measure_c2(args, rank, world, local)
                line["c3_f64"] = measure_c3(args, "f64", rank, world, local)
                line["c3_f32"] = measure_c3(args, "f32", rank, world, local)
                line["c5a"] = measure_c5(args, "c5a")
                line["c5b"] = measure_c5(args, "c5b")
            if not args.no_cpu:
                g, tg, ts, sample = cpu_sample_c2(args.ref_n)
                line["cpu_baseline"] = {"value": g, "unit": "GFLOP/s", "cores": 1, "kind": "port", "sample": sample, "host_cores_available": os.cpu_count()}
            else:
                line["cpu_baseline"] = None
        else:
            line = run_c4(args, rank, world, local)
            if not args.no_extras:
                line["c3_f64"] = measure_c3(args, "f64", rank, world, local, cpu=False)
                line["c3_f32"] = measure_c3(args, "f32", rank, world, local, cpu=False)
        if rank == 0:
            print(json.dumps(line), flush=True)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
