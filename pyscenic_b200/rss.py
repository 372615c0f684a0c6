# -*- coding: utf-8 -*-
"""
Regulon specificity scores on the GPU: drop-in for ``pyscenic.rss.regulon_specificity_scores``
(reference ``src/pyscenic/rss.py:8-33``), the Jensen-Shannon reduction of the AUC matrix per (cell type, regulon).

The reference loops over regulons x cell types and calls ``scipy.spatial.distance.jensenshannon`` on two n_cells
vectors each time (O(n_types * n_regulons * n_cells)); ``aucell_rss_f64`` (``csrc/aucell_rss.cu``) makes one pass over
the matrix, because a cell only enters the sum of its own type.
"""
from __future__ import annotations

import ctypes

import numpy as np
import pandas as pd

from . import _native

__all__ = ["regulon_specificity_scores"]


def regulon_specificity_scores(auc_mtx: pd.DataFrame, cell_type_series: pd.Series, *, device: int = 0) -> pd.DataFrame:
    """
    Calculates the Regulon Specificity Scores (RSS). [doi: 10.1016/j.celrep.2018.10.045]

    :param auc_mtx: The dataframe with the AUC values for all cells and regulons (n_cells x n_regulons).
    :param cell_type_series: A pandas Series object with cell identifiers as index and cell type labels as values
        (used positionally, as the reference does: ``(cell_type_series == cell_type)`` is paired with the rows of
        ``auc_mtx`` in order).
    :return: A pandas dataframe with the RSS values (cell type x regulon), float32 like the reference's.
    """
    import torch

    lib = _native.load()
    cell_types = list(cell_type_series.unique())                  # order of first appearance, as the reference
    regulons = list(auc_mtx.columns)
    n_cells, n_reg = auc_mtx.shape
    assert len(cell_type_series) == n_cells, "one cell type label per row of the AUC matrix"
    # a missing label is a "type" of its own in the reference's loop whose indicator (series == NaN) is all False:
    # its row comes out NaN, and its cells still count in every regulon's total.  Here they form a dummy type whose
    # row is dropped.
    valid_types = [t for t in cell_types if not pd.isna(t)]
    codes = pd.Categorical(cell_type_series.to_numpy(), categories=valid_types).codes.astype(np.int32)
    n_missing = int((codes < 0).sum())
    n_types = len(valid_types) + (1 if n_missing else 0)
    codes[codes < 0] = len(valid_types)
    if n_types == 0 or n_cells == 0:
        return pd.DataFrame(data=np.empty((len(cell_types), n_reg), dtype=np.float32), index=cell_types, columns=regulons)
    counts = np.bincount(codes, minlength=n_types).astype(np.int64)
    dev = torch.device("cuda", device)
    a = torch.from_numpy(np.ascontiguousarray(auc_mtx.to_numpy(), dtype=np.float64)).to(dev)
    lab = torch.from_numpy(codes).to(dev)
    out = torch.empty((n_types, n_reg), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        _native.check(lib.aucell_rss_f64(device, a.data_ptr(), n_cells, n_reg, n_reg, lab.data_ptr(), n_types,
                                         counts.ctypes.data_as(ctypes.c_void_p), out.data_ptr(), ctypes.c_void_p(st)))
    res = out.cpu().numpy().astype(np.float32)
    rows = np.full((len(cell_types), n_reg), np.nan, dtype=np.float32)
    k = 0
    for i, t in enumerate(cell_types):
        if not pd.isna(t):
            rows[i] = res[k]
            k += 1
    return pd.DataFrame(data=rows, index=cell_types, columns=regulons)
