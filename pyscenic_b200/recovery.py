# -*- coding: utf-8 -*-
"""
GPU versions of the ``ctxcore.recovery`` functions on the AUCell path (third-party dependency of the reference; call
sites ``src/pyscenic/aucell.py:15,95,122-126`` and ``src/pyscenic/transform.py:195``):

    derive_rank_cutoff(auc_threshold, total_genes, rank_threshold=None)
    aucs(rnk, total_genes, weights, auc_threshold)        AUC of ONE weighted gene set for every row of a ranking frame
    enrichment4cells(rnk_mtx, regulon, auc_threshold)     (= pyscenic.aucell.enrichment)

and the batched form of the cisTarget enrichment step (``src/pyscenic/transform.py:175-199``: ``db.load(module)`` ->
``calc_aucs`` -> NES -> threshold), all modules in ONE pass over the ranking database:

    module_aucs(rankings, modules, total_genes, auc_threshold, weighted_recovery)   features x modules AUC frame
    nes(aucs)                                                                       (auc - mean) / std per module
    enriched_features(aucs, nes_threshold)                                          (module, feature) -> NES, AUC

``aucs`` is the arithmetic cisTarget shares with AUCell (``transform.py:195`` calls it on a features x genes ranking
database restricted to a module's genes): the rows are motifs instead of cells and there is no ranking step, so it maps
onto the same scatter kernel (``aucell_from_rankings_u32``).
"""
from __future__ import annotations

import numpy as np
import pandas as pd

from .aucell import _torch_device, enrichment4cells  # noqa: F401  (re-exported)
from .plan import AUCellPlan, RegulonTables, derive_rank_cutoff  # noqa: F401  (re-exported)

__all__ = ["derive_rank_cutoff", "aucs", "enrichment4cells", "module_aucs", "nes", "enriched_features"]


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


def module_aucs(rankings: pd.DataFrame, modules, total_genes: int, auc_threshold: float,
                weighted_recovery: bool = False, *, rank_threshold=None, device: int = 0) -> pd.DataFrame:
    """
    AUC of every module (gene signature) for every feature (motif / track) of a ranking database.

    The reference does this one module at a time (``transform.py:175-195``: ``df = db.load(module)``, the database
    columns of the module's genes; ``weights`` = the module's weights or ones; ``calc_aucs(df, db.total_genes,
    weights, auc_threshold)``).  Here the database is uploaded once and one launch of the AUCell scatter kernel
    (``aucell_from_rankings_u32``) walks every feature row for all modules: feature := cell, module := regulon.

    :param rankings: features x genes frame of 0-based ranks (what ``db.load_full()`` returns; int16/int32).
    :param modules: signatures (``name``, ``genes``, ``sig[gene]``); genes absent from the database are ignored, as
        ``db.load`` does.  A module without any gene in the database gets a NaN column (the reference skips it).
    :param total_genes: ``db.total_genes``.
    :return: float64 frame, index = features, columns = module names.
    """
    import torch

    rank_cutoff = derive_rank_cutoff(auc_threshold, total_genes, rank_threshold)
    labels = np.asarray(rankings.columns, dtype=object)
    col_of = {}
    for j, g in enumerate(labels):
        col_of.setdefault(g, []).append(j)
    names, indptr, cols, wts, max_auc, valid = [], [0], [], [], [], []
    for m in modules:
        sel, w = [], []
        for g in m.genes:
            for j in col_of.get(g, ()):
                sel.append(j)
                w.append(float(m[g]) if weighted_recovery else 1.0)
        order = np.argsort(np.asarray(sel, dtype=np.int64), kind="stable")     # database column order, like db.load
        sel = np.asarray(sel, dtype=np.int32)[order]
        w = np.asarray(w, dtype=np.float64)[order]
        ok = len(sel) > 0
        names.append(m.name)
        valid.append(1 if ok else 0)
        max_auc.append(float((rank_cutoff + 1) * w.sum()) if ok else 0.0)
        if ok:
            assert max_auc[-1] > 0
            cols.append(sel)
            wts.append(w)
        indptr.append(indptr[-1] + (len(sel) if ok else 0))
    n_genes = len(labels)
    all_w = np.concatenate(wts) if wts else np.zeros(0)
    tables = RegulonTables(
        names=names, indptr=np.asarray(indptr, dtype=np.int64),
        cols=np.concatenate(cols).astype(np.int32) if cols else np.zeros(0, np.int32),
        weights=None if bool(np.all(all_w == 1.0)) else all_w, max_auc=np.asarray(max_auc, dtype=np.float64),
        valid=np.asarray(valid, dtype=np.uint8), rank_cutoff=int(rank_cutoff), n_genes=n_genes,
    )
    ranks = np.ascontiguousarray(rankings.to_numpy()).astype(np.uint32)
    dev = _torch_device(device)
    n_rows = ranks.shape[0]
    out = np.empty((n_rows, len(names)), dtype=np.float64)
    with AUCellPlan(n_genes, None, tables, device=device) as plan:
        chunk = max(1, (256 << 20) // (4 * max(n_genes, 1)))
        for s in range(0, n_rows, chunk):
            t = torch.from_numpy(ranks[s : s + chunk].view(np.int32)).to(dev)
            out[s : s + chunk] = plan.from_rankings(t).cpu().numpy()
    out[:, np.asarray(valid) == 0] = np.nan
    return pd.DataFrame(out, index=rankings.index, columns=pd.Index(names))


def nes(aucs):
    """Normalised enrichment score per module: ``(aucs - aucs.mean()) / aucs.std()`` over the features
    (``transform.py:196``; numpy's population standard deviation)."""
    a = aucs.to_numpy() if isinstance(aucs, pd.DataFrame) else np.asarray(aucs, dtype=np.float64)
    z = (a - a.mean(axis=0)) / a.std(axis=0)
    return pd.DataFrame(z, index=aucs.index, columns=aucs.columns) if isinstance(aucs, pd.DataFrame) else z


def enriched_features(aucs: pd.DataFrame, nes_threshold: float = 3.0) -> pd.DataFrame:
    """Features with ``NES >= nes_threshold`` per module (``transform.py:199-213``): frame indexed by
    (module, feature) with columns ``NES`` and ``AUC``."""
    z = nes(aucs)
    rows = []
    for name in aucs.columns:
        keep = z[name].to_numpy() >= nes_threshold
        for f, n_, a_ in zip(aucs.index[keep], z[name].to_numpy()[keep], aucs[name].to_numpy()[keep]):
            rows.append((name, f, n_, a_))
    idx = pd.MultiIndex.from_tuples([(r[0], r[1]) for r in rows], names=["module", "feature"])
    return pd.DataFrame({"NES": [r[2] for r in rows], "AUC": [r[3] for r in rows]}, index=idx)
