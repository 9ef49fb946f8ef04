"""Per-rank timeline of the distributed LU (torchrun, >= 2 GPUs): where each block step's time goes.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
        tools/mg_timeline.py [--order 65536] [--block 256] > gpurun_out/mg_timeline.jsonl

Every rank records 8 CUDA events per block step (lair_b200_mg_timeline); rank 0 gathers them and prints, per block
step, for the main stream of EVERY rank how long it waited for the panel of that step (the exposed part of the
owner's chain update -> panel -> pack -> broadcast), and for the owner the durations of the links of that chain.
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lair_b200 import _ffi, multigpu, sharding  # noqa: E402

P_READY, P_NEXT, P_PSTART, P_PDONE, P_PACK, P_BCAST, P_UPD, P_STEP = range(8)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--order", type=int, default=65536)
    ap.add_argument("--block", type=int, default=256)
    ap.add_argument("--every", type=int, default=8, help="print every k-th block step in full")
    args = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    L = _ffi.lib()
    multigpu.init()
    n, nb = args.order, args.block
    lcols = sharding.local_cols(n, nb, rank, world)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(4 + rank)
    a0 = torch.rand(n, lcols, dtype=torch.float64, device="cuda", generator=gen) * 10
    a = torch.empty_like(a0)
    for _ in range(2):  # warm-up
        a.copy_(a0)
        multigpu.getrf_mg(a, n, nb)
    torch.cuda.synchronize()
    dist.barrier()
    _ffi.check(L.lair_b200_mg_timeline(1))
    a.copy_(a0)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    multigpu.getrf_mg(a, n, nb)
    e1.record()
    torch.cuda.synchronize()
    total_ms = e0.elapsed_time(e1)
    nblk = sharding.num_blocks(n, nb)
    buf = np.zeros(nblk * 8, dtype=np.float32)
    got = ctypes.c_int64(0)
    _ffi.check(L.lair_b200_mg_timeline_read(buf.ctypes.data, buf.size, ctypes.byref(got)))
    _ffi.check(L.lair_b200_mg_timeline(0))
    t = torch.from_numpy(buf.reshape(nblk, 8)).cuda()
    allt = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allt, t)
    tot = torch.tensor([total_ms], device="cuda")
    allm = [torch.empty_like(tot) for _ in range(world)]
    dist.all_gather(allm, tot)
    if rank == 0:
        T = torch.stack(allt).cpu().numpy()  # [rank, block, point]
        wait = np.zeros((world, nblk))
        upd = np.zeros((world, nblk))
        for r in range(world):
            prev_done = np.concatenate([[0.0], T[r, :-1, P_STEP]])
            wait[r] = T[r, :, P_READY] - prev_done          # main stream idle until the panel of this step had landed
            upd[r] = T[r, :, P_STEP] - T[r, :, P_READY]     # update + left interchanges of this step
        owner = [(b + 1) % world for b in range(nblk)]      # owner of the NEXT block: its chain is recorded under step b
        chain = []
        for b in range(nblk - 1):
            o = owner[b]
            chain.append({"next_update_ms": float(T[o, b, P_NEXT] - T[o, b, P_READY]),
                          "panel_wait_ms": float(T[o, b, P_PSTART] - T[o, b, P_NEXT]),
                          "panel_ms": float(T[o, b, P_PDONE] - T[o, b, P_PSTART]),
                          "pack_ms": float(T[o, b, P_PACK] - T[o, b, P_PDONE]),
                          "bcast_ms": float(T[o, b, P_BCAST] - T[o, b, P_PACK])})
        print(json.dumps({"timeline": "getrf_mg", "n": n, "nb": nb, "world": world, "total_ms_per_rank": [float(m.item()) for m in allm],
                          "tflops": 2 / 3 * n ** 3 / max(float(m.item()) for m in allm) * 1e-9,
                          "main_stream_wait_ms_sum_per_rank": wait.sum(axis=1).round(2).tolist(),
                          "main_stream_update_ms_sum_per_rank": upd.sum(axis=1).round(2).tolist(),
                          "owner_chain_ms_sum": {k: round(sum(c[k] for c in chain if np.isfinite(c[k])), 2) for k in chain[0]}}), flush=True)
        for b in range(0, nblk - 1, args.every):
            print(json.dumps({"block": b, "rows_left": n - b * nb, "wait_ms_max_over_ranks": round(float(wait[:, b].max()), 3),
                              "wait_ms_mean": round(float(wait[:, b].mean()), 3), "update_ms_mean": round(float(upd[:, b].mean()), 3),
                              "update_ms_owner_of_next": round(float(upd[owner[b], b]), 3),
                              **{k: round(v, 3) for k, v in chain[b].items()}}), flush=True)
    multigpu.finalize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
