# -*- coding: utf-8 -*-
"""
GPU versions of the ``ctxcore.recovery`` functions on the AUCell path (third-party dependency of the reference; call
sites ``src/pyscenic/aucell.py:15,95,122-126`` and ``src/pyscenic/transform.py:195``):

    derive_rank_cutoff(auc_threshold, total_genes, rank_threshold=None)
    aucs(rnk, total_genes, weights, auc_threshold)        AUC of ONE weighted gene set for every row of a ranking frame
    enrichment4cells(rnk_mtx, regulon, auc_threshold)     (= pyscenic.aucell.enrichment)

``aucs`` is the arithmetic cisTarget shares with AUCell (``transform.py:195`` calls it on a features x genes ranking
database restricted to a module's genes): the rows are motifs instead of cells and there is no ranking step, so it maps
onto the same scatter kernel (``aucell_from_rankings_u32``).
"""
from __future__ import annotations

import numpy as np
import pandas as pd

from .aucell import _torch_device, enrichment4cells  # noqa: F401  (re-exported)
from .plan import AUCellPlan, RegulonTables, derive_rank_cutoff  # noqa: F401  (re-exported)

__all__ = ["derive_rank_cutoff", "aucs", "enrichment4cells"]


def aucs(rnk: pd.DataFrame, total_genes: int, weights: np.ndarray, auc_threshold: float, *, device: int = 0) -> np.ndarray:
    """
    AUC of one gene set (all columns of ``rnk``, with ``weights`` in column order) for every row of ``rnk``.

    :param rnk: ranking frame (n_features x n_selected_genes), 0-based ranks within the whole genome.
    :param total_genes: number of genes the ranks refer to (sets the rank cutoff).
    :param weights: weight of every column of ``rnk``.
    :param auc_threshold: fraction of the ranked genome to take into account.
    :return: float64 array (n_features,).
    """
    import torch

    rank_cutoff = derive_rank_cutoff(auc_threshold, total_genes)
    weights = np.ascontiguousarray(weights, dtype=np.float64)
    n_rows, n_cols = rnk.shape
    assert len(weights) == n_cols
    y_max = weights.sum()
    maxauc = float((rank_cutoff + 1) * y_max)
    assert maxauc > 0
    tables = RegulonTables(
        names=["aucs"], indptr=np.asarray([0, n_cols], dtype=np.int64), cols=np.arange(n_cols, dtype=np.int32),
        weights=None if bool(np.all(weights == 1.0)) else weights, max_auc=np.asarray([maxauc]),
        valid=np.ones(1, dtype=np.uint8), rank_cutoff=int(rank_cutoff), n_genes=n_cols,
    )
    ranks = np.ascontiguousarray(rnk.to_numpy()).astype(np.uint32)
    dev = _torch_device(device)
    out = np.empty((n_rows,), dtype=np.float64)
    with AUCellPlan(n_cols, None, tables, device=device) as plan:
        chunk = max(1, (256 << 20) // (4 * max(n_cols, 1)))
        for s in range(0, n_rows, chunk):
            t = torch.from_numpy(ranks[s : s + chunk].view(np.int32)).to(dev)
            out[s : s + chunk] = plan.from_rankings(t).cpu().numpy()[:, 0]
    return out
