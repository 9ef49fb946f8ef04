"""CPU, world_size 2 over gloo: the N>1 host logic (batch sharding, block-cyclic index maps)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from lair_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        full = rng.uniform(0, 10, size=(batch, 32, 32))
        start, count = sharding.batch_slice(batch, rank, world)
        mine = full[start:start + count].copy()
        piv, info = oracle.getrf_batched(mine)  # stands in for the per-rank GPU call; no collective on the data path
        # the only exchange is gathering results for the check
        sizes = [sharding.batch_slice(batch, r, world)[1] for r in range(world)]
        gathered = [torch.zeros(sizes[r], 32, 32, dtype=torch.float64) for r in range(world)]
        # gloo all_gather needs equal shapes: pad to the maximum
        mx = max(sizes)
        pad = torch.zeros(mx, 32, 32, dtype=torch.float64)
        pad[:count] = torch.from_numpy(mine)
        bufs = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad)
        if rank == 0:
            got = np.concatenate([bufs[r][:sizes[r]].numpy() for r in range(world)])
            ref = full.copy()
            oracle.getrf_batched(ref)
            np.save(os.path.join(out_dir, "ok.npy"), np.array([np.array_equal(got, ref)]))
    finally:
        dist.destroy_process_group()


def test_batch_sharding_two_ranks(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 101, str(tmp_path)), nprocs=2, join=True)
    assert np.load(tmp_path / "ok.npy")[0]


@pytest.mark.parametrize("batch,world", [(10, 1), (10, 3), (1_000_000, 8), (5, 8), (0, 2)])
def test_batch_slice_partitions(batch, world):
    slices = [sharding.batch_slice(batch, r, world) for r in range(world)]
    assert slices[0][0] == 0 and sum(c for _, c in slices) == batch
    for (s0, c0), (s1, _) in zip(slices, slices[1:]):
        assert s0 + c0 == s1
    assert max(c for _, c in slices) - min(c for _, c in slices) <= 1


@pytest.mark.parametrize("n,nb,world", [(65536, 512, 8), (1000, 128, 2), (8192, 256, 4), (300, 128, 8), (129, 128, 3)])
def test_block_cyclic_maps(n, nb, world):
    seen = np.zeros(n, dtype=int)
    for rank in range(world):
        lc = sharding.local_cols(n, nb, rank, world)
        for l in range(lc):
            g = sharding.local_to_global_col(l, nb, rank, world)
            assert g < n and sharding.global_to_local_col(g, nb, world) == (rank, l)
            seen[g] += 1
    assert np.all(seen == 1)
    for col in (0, 1, nb - 1, nb, n // 2, n - 1, n):
        for rank in range(world):
            exp = sum(1 for l in range(sharding.local_cols(n, nb, rank, world))
                      if sharding.local_to_global_col(l, nb, rank, world) < col)
            assert sharding.first_local_col_at_or_after(col, nb, rank, world) == exp
