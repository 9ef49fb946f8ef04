"""Index arithmetic of the two multi-GPU decompositions of the LU path.

* batched small LU: contiguous slices of the batch, one per rank, no data-path collective;
* one large LU: 1-D block-cyclic COLUMN distribution (block j of width nb lives on rank j mod P),
  so pivot search, panel and row interchanges are local to the owner of a block column and the
  only exchange is the broadcast of the factored panel + its pivots.
Pure functions, used by bench.py / the multi-GPU driver and covered by 2-rank gloo tests on CPU.
"""
from __future__ import annotations


def batch_slice(batch: int, rank: int, world: int) -> tuple[int, int]:
    """(start, count) of rank's contiguous share of `batch` items; shares differ by at most one."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, rem = divmod(batch, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def block_owner(block: int, world: int) -> int:
    return block % world


def num_blocks(n: int, nb: int) -> int:
    return (n + nb - 1) // nb


def local_blocks(n: int, nb: int, rank: int, world: int) -> list[int]:
    """Global block-column indices stored on `rank`, in local order."""
    return list(range(rank, num_blocks(n, nb), world))


def local_cols(n: int, nb: int, rank: int, world: int) -> int:
    """Number of matrix columns stored on `rank`."""
    return sum(min(nb, n - b * nb) for b in local_blocks(n, nb, rank, world))


def global_to_local_col(col: int, nb: int, world: int) -> tuple[int, int]:
    """(owner rank, local column index) of global column `col`."""
    block, off = divmod(col, nb)
    return block % world, (block // world) * nb + off


def local_to_global_col(lcol: int, nb: int, rank: int, world: int) -> int:
    lblock, off = divmod(lcol, nb)
    return (lblock * world + rank) * nb + off


def first_local_col_at_or_after(col: int, nb: int, rank: int, world: int) -> int:
    """Local index of the first column stored on `rank` whose global index is >= `col`
    (== the number of local columns strictly left of global column `col`)."""
    block, off = divmod(col, nb)
    full_cycles, r = divmod(block, world)
    lcol = full_cycles * nb
    if rank < r:
        lcol += nb
    elif rank == r:
        lcol += off
    return lcol
