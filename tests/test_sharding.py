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
