# -*- coding: utf-8 -*-
"""
Cell-sharded AUCell across the GPUs of one box: one process per GPU (``torch.distributed``), contiguous cell ranges,
regulon tables replicated, NO collective on the data path -- every cell's ranks and AUCs depend only on that cell's
row.  The only communication is the final gather of the ``[cells_shard x regulons]`` float64 blocks
(and a ``[regulons]`` max-reduce when ``normalize=True``, reference ``aucell.py:182``).

The reference parallelises the other axis (regulon chunks across forked workers over a shared ranking matrix,
``aucell.py:151-166``); on GPUs sharding cells avoids replicating the big operand.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

__all__ = ["shard_bounds", "shard_bounds_by_nnz", "gather_blocks", "global_column_max", "aucell_sharded"]


def shard_bounds(n_cells: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous, near-equal cell ranges; rank r owns ``[bounds[r][0], bounds[r][1])`` (possibly empty)."""
    edges = [(n_cells * i) // world_size for i in range(world_size + 1)]
    return [(edges[i], edges[i + 1]) for i in range(world_size)]


def shard_bounds_by_nnz(indptr: np.ndarray, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous cell ranges holding near-equal numbers of stored entries (CSR): the kernel's work per cell is
    proportional to its non-zeros, so this balances time rather than cell counts."""
    indptr = np.asarray(indptr, dtype=np.int64)
    n_cells = len(indptr) - 1
    total = int(indptr[-1] - indptr[0])
    targets = indptr[0] + (np.arange(1, world_size, dtype=np.float64) * total / world_size)
    cuts = np.searchsorted(indptr, targets, side="left")
    edges = [0] + [int(min(max(c, 0), n_cells)) for c in cuts] + [n_cells]
    edges = list(np.maximum.accumulate(edges))
    return [(int(edges[i]), int(edges[i + 1])) for i in range(world_size)]


def gather_blocks(block, bounds: Sequence[Tuple[int, int]], n_cols: int, group=None, dst: int = 0):
    """Gather per-rank ``[rows_r x n_cols]`` float64 tensors onto ``dst`` in rank order.  Returns the full matrix on
    ``dst`` and ``None`` elsewhere.  Works with the ``nccl`` (CUDA tensors) and ``gloo`` (CPU tensors) backends."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    rows = [b[1] - b[0] for b in bounds]
    assert block.shape == (rows[rank], n_cols) and block.dtype == torch.float64
    pad = max(rows) if rows else 0
    send = torch.zeros((pad, n_cols), dtype=torch.float64, device=block.device)
    send[: rows[rank]] = block
    recv = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, recv, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([recv[r][: rows[r]] for r in range(world)], dim=0)


def global_column_max(block, group=None):
    """Per-regulon maximum over the cells of ALL ranks (for ``normalize=True``)."""
    import torch
    import torch.distributed as dist

    if block.shape[0]:
        m = block.max(dim=0).values
    else:
        m = torch.full((block.shape[1],), -float("inf"), dtype=block.dtype, device=block.device)
    dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    return m


def aucell_sharded(compute_shard: Callable[[int, int], "object"], n_cells: int, n_regulons: int,
                   bounds: Optional[Sequence[Tuple[int, int]]] = None, normalize: bool = False, group=None,
                   dst: int = 0):
    """Run ``compute_shard(c0, c1) -> float64 tensor [c1 - c0, n_regulons]`` on this rank's cell range and gather.

    ``compute_shard`` is where the GPU work happens (e.g. ``AUCellPlan.aucell_csr`` on the rank's slice); it is a
    parameter so that the partition / gather logic can be exercised with the ``gloo`` backend on CPUs."""
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if bounds is None:
        bounds = shard_bounds(n_cells, world)
    c0, c1 = bounds[rank]
    block = compute_shard(c0, c1)
    if normalize:
        block = block / global_column_max(block, group)
    return gather_blocks(block, bounds, n_regulons, group=group, dst=dst)
