# -*- coding: utf-8 -*-
"""Multi-GPU parity on hardware (``-m gpu``; skipped below two devices): the cell-sharded drivers must reproduce the
single-GPU result bit for bit -- ``aucell(..., devices=[...])`` in one process, and the
one-process-per-GPU driver (``pyscenic_b200.sharding``) over NCCL."""
import os
import socket
import subprocess
import sys

import numpy as np
import pandas as pd
import pytest

import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def test_in_process_devices_equal_single_gpu():
    import scipy.sparse as sp

    from pyscenic_b200.aucell import aucell

    n = _n_gpus()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    devs = list(range(min(n, 8)))
    rs = np.random.RandomState(0)
    n_cells, n_genes = 3001, 5000
    X = synth.dense_counts(rs, n_cells, n_genes, 0.1)
    sigs = synth.regulons_to_signatures(*synth.make_regulons(rs, n_genes, 40, 10, 300, weighted=True))
    cols = ["g%d" % j for j in range(n_genes)]
    frames = (("f32", pd.DataFrame(X, columns=cols)),
              ("f64", pd.DataFrame(X.astype(np.float64) + 1e-13 * (X > 0), columns=cols)),
              ("f32-cmajor", pd.DataFrame(np.ascontiguousarray(X), columns=cols, copy=False)))
    for name, df in frames:
        a = aucell(df, sigs, seed=3, devices=[0])
        b = aucell(df, sigs, seed=3, devices=devs)
        c = aucell(df, sigs, seed=3, devices=[devs[-1]])
        assert np.array_equal(a.values, b.values) and np.array_equal(a.values, c.values), name
    csr = sp.csr_matrix(X)
    a = aucell(csr, sigs, seed=3, genes=cols, devices=[0])
    b = aucell(csr, sigs, seed=3, genes=cols, devices=devs)
    assert np.array_equal(a.values, b.values)


def test_one_process_per_gpu_sharded_driver_is_bit_identical():
    """torchrun, one rank per GPU: partition ONE matrix by stored entries, per-rank kernel, NCCL gather to rank 0,
    compared there with the whole matrix computed on one GPU (scripts/sharded_check.py prints the verdict)."""
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(n, 8)),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "scripts", "sharded_check.py"),
           "--cells", "60000"]
    res = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:]
    assert '"identical": true' in res.stdout, res.stdout[-3000:]
    assert '"identical_shared_host_memory": true' in res.stdout, res.stdout[-3000:]      # result assembled in /dev/shm
