# -*- coding: utf-8 -*-
"""World-size-2 gloo tests (CPU) of the cell-sharding / gather logic used by the multi-GPU driver."""
import os
import socket

import numpy as np
import pytest

from pyscenic_b200.sharding import shard_bounds, shard_bounds_by_nnz


def test_shard_bounds_cover_all_cells():
    for n in (0, 1, 7, 1000, 1_300_000):
        for w in (1, 2, 4, 8):
            b = shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n and len(b) == w
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1


def test_shard_bounds_by_nnz_balances_entries():
    rs = np.random.RandomState(0)
    nnz = rs.lognormal(7.5, 0.5, size=5000).astype(np.int64)
    indptr = np.concatenate([[0], np.cumsum(nnz)])
    for w in (2, 4, 8):
        b = shard_bounds_by_nnz(indptr, w)
        assert b[0][0] == 0 and b[-1][1] == 5000
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        loads = [indptr[e] - indptr[s] for s, e in b]
        assert max(loads) / (indptr[-1] / w) < 1.05
    # degenerate: empty matrix, more ranks than cells
    assert shard_bounds_by_nnz(np.zeros(1, dtype=np.int64), 4) == [(0, 0)] * 4
    b = shard_bounds_by_nnz(np.array([0, 5, 9]), 4)
    assert b[0][0] == 0 and b[-1][1] == 2 and all(s <= e for s, e in b)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_cells, n_reg, normalize, out_path):
    import torch
    import torch.distributed as dist

    from pyscenic_b200.sharding import aucell_sharded

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.from_numpy(np.random.RandomState(3).rand(n_cells, n_reg))     # same on every rank

    def compute(c0, c1):            # stands in for the per-GPU kernel call on the rank's cell range
        return full[c0:c1].clone()

    res = aucell_sharded(compute, n_cells, n_reg, normalize=normalize)
    if rank == 0:
        want = full / full.max(dim=0).values if normalize else full
        ok = res is not None and res.shape == want.shape and torch.equal(res, want)
        open(out_path, "w").write("ok" if ok else "mismatch")
    else:
        assert res is None
    dist.destroy_process_group()


@pytest.mark.parametrize("n_cells,normalize", [(101, False), (101, True), (1, False)])
def test_gloo_world2_gather_matches_single_process(tmp_path, n_cells, normalize):
    import torch.multiprocessing as mp

    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, _free_port(), n_cells, 7, normalize, out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def _scatter_worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist

    from pyscenic_b200.sharding import gather_blocks, scatter_csr_shards

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    indptr = indices = data = None
    rs = np.random.RandomState(9)
    nnz = rs.randint(0, 40, size=57).astype(np.int64)
    ip = np.concatenate([[0], np.cumsum(nnz)])
    ix = rs.randint(0, 1000, size=int(ip[-1])).astype(np.int32)
    dv = rs.rand(int(ip[-1])).astype(np.float32)
    if rank == 0:                                   # only the source holds the matrix
        indptr, indices, data = torch.from_numpy(ip), torch.from_numpy(ix), torch.from_numpy(dv)
    bounds, s_ip, s_ix, s_dv = scatter_csr_shards(indptr, indices, data)
    c0, c1 = bounds[rank]
    ok = (bounds == shard_bounds_by_nnz(ip, world)
          and np.array_equal(s_ip.numpy(), ip[c0:c1 + 1])
          and np.array_equal(s_ix.numpy(), ix[ip[c0]:ip[c1]]) and np.array_equal(s_dv.numpy(), dv[ip[c0]:ip[c1]]))
    # a per-cell "kernel": the row sums; gathered blocks must reproduce the whole matrix' row sums in order
    rel = (s_ip - s_ip[0]).numpy()
    block = torch.tensor([[float(s_dv.numpy()[rel[i]:rel[i + 1]].astype(np.float64).sum())] for i in range(c1 - c0)],
                         dtype=torch.float64).reshape(c1 - c0, 1)
    full = gather_blocks(block, bounds, 1)
    if rank == 0:
        want = np.array([dv[ip[i]:ip[i + 1]].astype(np.float64).sum() for i in range(57)])
        ok = ok and np.array_equal(full.numpy()[:, 0], want)
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        open(out_path, "w").write("ok" if int(flag) == 1 else "mismatch")
    dist.destroy_process_group()


def test_gloo_world2_scatter_of_one_matrix_and_gather(tmp_path):
    """The strong-scaling driver: ONE CSR matrix on rank 0, partitioned by stored entries, every rank receives its own
    cell range (scatter_csr_shards), the per-rank blocks come back in order (gather_blocks)."""
    import torch.multiprocessing as mp

    out = str(tmp_path / "result.txt")
    mp.spawn(_scatter_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def _shared_result_worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist

    from pyscenic_b200 import sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_cells, n_reg = 103, 5
    bounds = sharding.shard_bounds(n_cells, world)
    c0, c1 = bounds[rank]
    full = sharding._shared_result(n_cells, n_reg, (c0, c1), None, 0)
    assert full.shape == (n_cells, n_reg)
    want = np.random.RandomState(5).rand(n_cells, n_reg)             # same on every rank
    full[c0:c1] = want[c0:c1]                                        # every rank writes its own cell range
    dist.barrier()
    ok = True
    if rank == 0:
        ok = np.array_equal(np.asarray(full), want)                  # rank 0 sees what the others wrote
    again = sharding._shared_result(n_cells, n_reg, (c0, c1), None, 0)
    ok = ok and again is full                                        # kept between calls of the same shape
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        leftovers = [f for f in os.listdir("/dev/shm") if f.startswith("aucell_b200_%d_" % os.getpid())]
        open(out_path, "w").write("ok" if int(flag) == 1 and not leftovers else "mismatch")
    dist.destroy_process_group()


@pytest.mark.skipif(not os.path.isdir("/dev/shm"), reason="needs /dev/shm")
def test_gloo_world2_result_in_shared_host_memory(tmp_path):
    """The result buffer of the one-box multi-GPU job (sharding._shared_result): created by rank 0 under /dev/shm, mapped
    by every rank, written in cell ranges, read by rank 0; unlinked at once, kept between calls."""
    import torch.multiprocessing as mp

    out = str(tmp_path / "result.txt")
    mp.spawn(_shared_result_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert open(out).read() == "ok"
