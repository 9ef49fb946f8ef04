"""Multi-GPU host layer (one process per GPU; `torch.distributed` is only the plumbing that carries
the NCCL unique id and the test/bench reductions -- the data path is the library's own
ncclBroadcast of factored panels, see lair_b200/csrc/mg.cu).

    init()                              -> create the library's NCCL communicator for this rank
    distribute_columns(a_full, nb)      -> this rank's block-cyclic column slab of a global matrix
    getrf_mg(a_local, n, nb)            -> in-place distributed LU; returns (ipiv, info) tensors
    gather_columns(a_local, n, nb)      -> reassemble the global matrix (tests only)
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import _ffi, sharding

_initialised = False


def init() -> None:
    """Create the NCCL communicator inside liblair_b200 from an id broadcast over torch.distributed."""
    global _initialised
    if _initialised:
        return
    rank, world = dist.get_rank(), dist.get_world_size()
    L = _ffi.lib()
    _ffi.check(L.lair_b200_init(torch.cuda.current_device()))
    buf = (ctypes.c_ubyte * 128)()
    if rank == 0:
        _ffi.check(L.lair_b200_mg_unique_id(buf))
    t = torch.tensor(list(buf), dtype=torch.uint8, device="cuda")
    dist.broadcast(t, src=0)
    raw = bytes(t.cpu().tolist())
    idbuf = (ctypes.c_ubyte * 128).from_buffer_copy(raw)
    _ffi.check(L.lair_b200_mg_init(rank, world, idbuf))
    _initialised = True


def finalize() -> None:
    global _initialised
    if _initialised:
        _ffi.check(_ffi.lib().lair_b200_mg_finalize())
        _initialised = False


def local_col_indices(n: int, nb: int, rank: int, world: int) -> torch.Tensor:
    cols = [sharding.local_to_global_col(l, nb, rank, world) for l in range(sharding.local_cols(n, nb, rank, world))]
    return torch.tensor(cols, dtype=torch.long)


def distribute_columns(a_full: torch.Tensor, nb: int) -> torch.Tensor:
    """This rank's columns of a replicated global matrix, as a contiguous (n, local_cols) tensor."""
    rank, world = dist.get_rank(), dist.get_world_size()
    idx = local_col_indices(a_full.shape[1], nb, rank, world).to(a_full.device)
    return a_full.index_select(1, idx).contiguous()


def getrf_mg(a_local: torch.Tensor, n: int, nb: int):
    """Distributed in-place LU of the n x n matrix whose block-cyclic column slab is `a_local`."""
    assert a_local.is_cuda and a_local.is_contiguous() and a_local.shape[0] == n
    pfx = {torch.float64: "d", torch.float32: "s"}[a_local.dtype]
    ipiv = torch.empty(n, dtype=torch.int32, device=a_local.device)
    info = torch.empty(1, dtype=torch.int32, device=a_local.device)
    fn = getattr(_ffi.lib(), f"lair_b200_{pfx}getrf_mg_dev")
    stream = torch.cuda.current_stream().cuda_stream
    _ffi.check(fn(n, nb, a_local.data_ptr(), a_local.shape[1], ipiv.data_ptr(), info.data_ptr(), stream))
    return ipiv, info


def gather_columns(a_local: torch.Tensor, n: int, nb: int) -> torch.Tensor:
    """All ranks' slabs reassembled into the global n x n matrix on every rank (tests only)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    full = torch.zeros(n, n, dtype=a_local.dtype, device=a_local.device)
    idx = local_col_indices(n, nb, rank, world).to(a_local.device)
    full.index_copy_(1, idx, a_local)
    dist.all_reduce(full)
    return full
