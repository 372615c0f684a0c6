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

__all__ = ["shard_bounds", "shard_bounds_by_nnz", "gather_blocks", "global_column_max", "aucell_sharded",
           "aucell_csr_shard_from_host", "aucell_csr_shard_to_shared_host", "scatter_csr_shards"]


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


def _all_on_one_host(group=None) -> bool:
    import socket

    import torch.distributed as dist

    names = [None] * dist.get_world_size(group)
    dist.all_gather_object(names, socket.gethostname(), group=group)
    return len(set(names)) == 1


_SHARED_RESULT = {}      # (n_cells, n_regulons, dst, group id) -> (memmap, registered) : the job's result buffer, kept


def _shared_result(n_cells: int, n_regulons: int, rows: Tuple[int, int], group, dst: int):
    """The ``[n_cells x n_regulons]`` float64 result of a one-box job in shared host memory: rank ``dst`` creates a file
    under /dev/shm, every rank maps it and page-locks the rows it will write (``cudaHostRegister``), so that its
    device-to-host copies land there at full PCIe rate.  Created on first use and kept for later calls of the same
    shape (like a pinned-memory pool: mapping, zero-filling and page-locking 5 GB costs about a second); the file name
    is unlinked as soon as every rank holds the mapping."""
    import os

    import torch
    import torch.distributed as dist

    key = (n_cells, n_regulons, dst, id(group))
    hit = _SHARED_RESULT.get(key)
    if hit is not None:
        return hit[0]
    if not _all_on_one_host(group):
        raise RuntimeError("the shared-host-memory result needs every rank on one box; use aucell_csr_shard_from_host")
    rank = dist.get_rank(group)
    name = [None]
    if rank == dst:
        name[0] = "/dev/shm/aucell_b200_%d_%s.f64" % (os.getpid(), os.urandom(6).hex())
        try:
            with open(name[0], "wb") as f:
                # reserve the pages now: a /dev/shm that is too small must fail here, not with SIGBUS in the middle of a copy
                os.posix_fallocate(f.fileno(), 0, max(1, n_cells * n_regulons * 8))
        except OSError as e:
            try:
                os.unlink(name[0])
            except OSError:
                pass
            name[0] = "!%s" % e
    dist.broadcast_object_list(name, src=dst, group=group)
    if name[0].startswith("!"):
        raise RuntimeError("no room for the result in /dev/shm (%s); use aucell_csr_shard_from_host" % name[0][1:])
    full = np.memmap(name[0], dtype=np.float64, mode="r+", shape=(n_cells, n_regulons))
    registered = False
    c0, c1 = rows
    if (c1 > c0 or rank == dst) and torch.cuda.is_available():
        # `dst` page-locks the whole matrix (a host range that is only partly registered cannot be the source of one
        # cudaMemcpy: whoever consumes the result would trip over it), the other ranks the rows they write
        base = full.ctypes.data
        end = (base + n_cells * n_regulons * 8 + 4095) & ~4095
        b = base if rank == dst else (base + c0 * n_regulons * 8) & ~4095
        e = end if rank == dst else min(end, (base + c1 * n_regulons * 8 + 4095) & ~4095)
        registered = int(torch.cuda.cudart().cudaHostRegister(b, e - b, 0)) == 0
    dist.barrier(group=group)
    if rank == dst:
        os.unlink(name[0])
    _SHARED_RESULT[key] = (full, registered)
    return full


def aucell_csr_shard_to_shared_host(plan, h_indptr, h_indices, h_data, bounds: Sequence[Tuple[int, int]], group=None,
                                    dst: int = 0, timings: Optional[dict] = None):
    """The cell-sharded job on ONE box without any device-side gather: the ``[cells x regulons]`` float64 result lives in
    shared host memory (``_shared_result``), and every rank lets the library's pipelined host call
    (``aucell_csr_f32_host``: upload, kernels and download of consecutive chunks overlap) write its own cell range
    straight into it -- each over its own PCIe link.  What is left of the "gather" is a barrier.  Returns the full
    matrix on ``dst`` (a numpy array over the mapping, valid until the next call of the same shape) and ``None``
    elsewhere.  ``h_indptr`` / ``h_indices`` / ``h_data``: this rank's cell range as torch CPU tensors or numpy arrays
    (pinned for full PCIe rate)."""
    import time

    import torch.distributed as dist

    rank = dist.get_rank(group)
    n_cells = bounds[-1][1]
    R = plan.n_regulons
    c0, c1 = bounds[rank]
    t0 = time.perf_counter()
    full = _shared_result(n_cells, R, (c0, c1), group, dst)
    t1 = time.perf_counter()
    if c1 > c0:
        as_np = lambda t: t.numpy() if hasattr(t, "numpy") else np.asarray(t)
        ip = np.ascontiguousarray(as_np(h_indptr), dtype=np.int64)
        assert len(ip) == c1 - c0 + 1
        ip = ip - ip[0]                       # (the shard's own entries start at 0)
        plan.aucell_csr_host(ip, as_np(h_indices), as_np(h_data), out=full[c0:c1])
    t2 = time.perf_counter()
    dist.barrier(group=group)
    t3 = time.perf_counter()
    if timings is not None:
        timings.update(setup_s=t1 - t0, upload_kernel_download_s=t2 - t1, barrier_s=t3 - t2)
    return full if rank == dst else None


def aucell_csr_shard_from_host(plan, h_indptr, h_indices, h_data, bounds: Sequence[Tuple[int, int]], group=None,
                               dst: int = 0, to_host: bool = True, timings: Optional[dict] = None):
    """One rank's part of a cell-sharded AUCell job whose CSR matrix lives in HOST memory (the multi-GPU form of
    ``pyscenic.aucell.aucell``, reference ``src/pyscenic/aucell.py:131-181``, which shares one ranking matrix between
    forked workers; here cells are partitioned and only the result travels).

    ``h_indptr`` / ``h_indices`` / ``h_data`` are torch CPU tensors (pinned for full PCIe rate) holding THIS rank's
    cell range ``bounds[rank]``: ``h_indptr`` has ``c1 - c0 + 1`` entries (absolute or rebased offsets), the other two
    the entries of those cells.  Upload, kernel, NCCL gather of the ``[cells x regulons]`` blocks to ``dst``, and (when
    ``to_host``) the copy of the gathered matrix into pinned host memory there.  Returns the full matrix on ``dst``
    (a pinned CPU tensor, or the CUDA tensor when ``to_host`` is false) and ``None`` on the other ranks.
    ``timings`` (optional dict) receives the wall-clock seconds of the phases, measured around device synchronisations.
    """
    import time

    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    dev = torch.device("cuda", plan.device)
    c0, c1 = bounds[rank]
    assert h_indptr.numel() == c1 - c0 + 1
    t0 = time.perf_counter()
    d_ip = h_indptr.to(dev, non_blocking=True)
    d_ix = h_indices.to(dev, non_blocking=True)
    d_dv = h_data.to(dev, non_blocking=True)
    d_ip = d_ip - d_ip[0]
    block = plan.aucell_csr(d_ip, d_ix, d_dv)
    torch.cuda.synchronize(dev)
    t1 = time.perf_counter()
    full = gather_blocks(block, bounds, plan.n_regulons, group=group, dst=dst)
    torch.cuda.synchronize(dev)
    t2 = time.perf_counter()
    res = None
    if rank == dst:
        if to_host:
            res = torch.empty(full.shape, dtype=full.dtype, pin_memory=True)
            res.copy_(full, non_blocking=True)
            torch.cuda.synchronize(dev)
        else:
            res = full
    t3 = time.perf_counter()
    if timings is not None:
        timings.update(upload_and_kernel_s=t1 - t0, gather_s=t2 - t1, to_host_s=t3 - t2)
    return res


def scatter_csr_shards(indptr, indices, data, world_size: Optional[int] = None, group=None, src: int = 0):
    """Distribute ONE CSR matrix that lives on ``src`` (torch tensors on its device; ``None`` on the other ranks): the
    partition by stored entries is made on ``src`` and every rank receives its own cell range.  Returns
    ``(bounds, indptr_shard, indices_shard, data_shard)`` -- tensors on the rank's current device (``nccl``) or on the
    CPU (``gloo``), ``indptr_shard`` holding the absolute offsets of ``bounds[rank]``."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    world_size = world if world_size is None else world_size
    meta = [None]
    if rank == src:
        ip = indptr.cpu().numpy()
        bounds = shard_bounds_by_nnz(ip, world_size)
        meta = [(bounds, [(int(ip[a]), int(ip[b])) for a, b in bounds], str(indices.dtype), str(data.dtype))]
    dist.broadcast_object_list(meta, src=src, group=group)
    bounds, spans, ix_dt, dv_dt = meta[0]
    if rank == src:
        for r in range(world):
            if r == src:
                continue
            (c0, c1), (e0, e1) = bounds[r], spans[r]
            dist.send(indptr[c0:c1 + 1].contiguous(), dst=r, group=group)
            if e1 > e0:
                dist.send(indices[e0:e1].contiguous(), dst=r, group=group)
                dist.send(data[e0:e1].contiguous(), dst=r, group=group)
        (c0, c1), (e0, e1) = bounds[src], spans[src]
        return bounds, indptr[c0:c1 + 1], indices[e0:e1], data[e0:e1]
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    (c0, c1), (e0, e1) = bounds[rank], spans[rank]
    s_ip = torch.empty(c1 - c0 + 1, dtype=torch.int64, device=dev)
    s_ix = torch.empty(e1 - e0, dtype=getattr(torch, ix_dt.split(".")[-1]), device=dev)
    s_dv = torch.empty(e1 - e0, dtype=getattr(torch, dv_dt.split(".")[-1]), device=dev)
    dist.recv(s_ip, src=src, group=group)
    if e1 > e0:
        dist.recv(s_ix, src=src, group=group)
        dist.recv(s_dv, src=src, group=group)
    return bounds, s_ip, s_ix, s_dv
