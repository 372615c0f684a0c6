# -*- coding: utf-8 -*-
"""
Binarization of an AUC matrix on the GPU: drop-in for ``pyscenic.binarization`` (reference
``src/pyscenic/binarization.py:12-98``): the same ``derive_threshold`` / ``binarize`` signatures and results, computed
for all regulons at once by the kernels of ``csrc/aucell_binarize.cu`` instead of one regulon at a time in a process
pool.

Per regulon the reference (a) decides whether the AUC distribution is bimodal -- ``method="hdt"``: Hartigan's dip
test, p <= 0.05; ``method="bic"``: BIC of a two-component against a one-component Gaussian mixture --, (b) if not:
``mean + 2 * std``, (c) if so: the minimum of the Gaussian kernel density (``scipy.stats.gaussian_kde``, Scott's rule)
between the two means of a two-component mixture, found with scipy's bounded Brent minimiser (``xatol=1e-5``).

Differences to the reference, all inside its own tolerances:

* the EM starts from the optimal two-means split of the sorted values instead of a seeded k-means++ run (``seed`` is
  accepted and ignored: the result is deterministic); the mixture means only bound the minimiser;
* the dip test's p-value is read from a table of the dip distribution of uniform samples computed with THIS
  implementation of the statistic (``pyscenic_b200/dip_table.json``, written by ``scripts/make_dip_table.py``), not from
  the ``diptest`` package's table: ``diptest`` is not installed here, so the dip decision is not pinned against it.
"""
from __future__ import annotations

import ctypes
import json
import os
from typing import Mapping, Optional

import numpy as np
import pandas as pd

from . import _native

__all__ = ["derive_threshold", "derive_thresholds", "binarize", "dip_statistics", "dip_pvalues"]

_DIP_TABLE = None


def _dip_table():
    global _DIP_TABLE
    if _DIP_TABLE is None:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "dip_table.json")) as fh:
            t = json.load(fh)
        _DIP_TABLE = (np.asarray(t["n"], dtype=np.float64), np.asarray(t["prob"], dtype=np.float64),
                      np.asarray(t["sqrt_n_dip"], dtype=np.float64))
    return _DIP_TABLE


def dip_pvalues(dips: np.ndarray, n: int) -> np.ndarray:
    """p-value of the dip test for samples of size ``n``: P(dip of a uniform sample >= observed), by interpolation of
    ``sqrt(n) * dip`` between the table's sample sizes and probabilities (the scheme of the ``diptest`` package)."""
    ns, probs, crit = _dip_table()
    nn = float(min(max(n, ns[0]), ns[-1]))
    i1 = int(np.searchsorted(ns, nn, side="left"))
    i0 = max(i1 - 1, 0)
    i1 = min(i1, len(ns) - 1)
    f = 0.0 if ns[i1] == ns[i0] else (nn - ns[i0]) / (ns[i1] - ns[i0])
    y = crit[i0] + f * (crit[i1] - crit[i0])
    sd = np.sqrt(float(n)) * np.asarray(dips, dtype=np.float64)
    return 1.0 - np.interp(sd, y, probs, left=0.0, right=1.0)


def _sorted_rows(auc_mtx: pd.DataFrame, device: int):
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("pyscenic_b200 needs a CUDA device (sm_100); there is no CPU fallback")
    dev = torch.device("cuda", device)
    a = torch.from_numpy(np.ascontiguousarray(auc_mtx.to_numpy(), dtype=np.float64)).to(dev)
    cols = a.t().contiguous()                       # [regulons x cells]
    return torch.sort(cols, dim=1).values.contiguous(), dev


def _stream(dev):
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def dip_statistics(sorted_rows, device: int = 0):
    """Hartigan's dip statistic of every (ascending) row of a CUDA fp64 tensor ``[regulons x cells]``."""
    import torch

    lib = _native.load()
    n_reg, n = sorted_rows.shape
    out = torch.empty(n_reg, dtype=torch.float64, device=sorted_rows.device)
    batch = max(1, int((1 << 30) // (16 * (n + 1))))           # <= 1 GiB of scratch at a time
    for r0 in range(0, n_reg, batch):
        r1 = min(n_reg, r0 + batch)
        scratch = torch.empty((r1 - r0) * 4 * (n + 1), dtype=torch.int32, device=sorted_rows.device)
        _native.check(lib.aucell_binarize_dip_f64(device, sorted_rows[r0:r1].data_ptr(), r1 - r0, n, scratch.data_ptr(),
                                                  out[r0:r1].data_ptr(), _stream(sorted_rows.device)))
        torch.cuda.synchronize(sorted_rows.device)
    return out


def _gmm2(sorted_rows, device: int):
    """Two-component mixtures of every row; returns the [rows x 9] parameter block (see aucell_binarize_gmm2_f64)."""
    import torch

    lib = _native.load()
    n_reg, n = sorted_rows.shape
    # optimal two-means split of sorted values: maximise the between-class sum of squares over the split point
    cs = torch.cumsum(sorted_rows, dim=1)
    k = torch.arange(1, n, dtype=torch.float64, device=sorted_rows.device)
    if n > 1:
        left, tot = cs[:, :-1], cs[:, -1:]
        score = left * left / k + (tot - left) * (tot - left) / (float(n) - k)
        split = (torch.argmax(score, dim=1) + 1).to(torch.int32)
    else:
        split = torch.ones(n_reg, dtype=torch.int32, device=sorted_rows.device)
    out = torch.empty((n_reg, 9), dtype=torch.float64, device=sorted_rows.device)
    _native.check(lib.aucell_binarize_gmm2_f64(device, sorted_rows.data_ptr(), n_reg, n, split.data_ptr(), 100, 1e-3, 1e-6,
                                               out.data_ptr(), _stream(sorted_rows.device)))
    return out


def derive_thresholds(auc_mtx: pd.DataFrame, seed=None, method: str = "hdt", *, device: int = 0) -> pd.Series:
    """Thresholds of ALL regulons (the columns of ``auc_mtx``) in one pass over the GPU; what ``binarize`` uses."""
    import torch

    assert auc_mtx is not None and not auc_mtx.empty
    assert method in {"hdt", "bic"}
    lib = _native.load()
    rows, dev = _sorted_rows(auc_mtx, device)
    n_reg, n = rows.shape
    stats = torch.empty((n_reg, 4), dtype=torch.float64, device=dev)
    _native.check(lib.aucell_binarize_stats_f64(device, rows.data_ptr(), n_reg, n, stats.data_ptr(), _stream(dev)))
    stats_h = stats.cpu().numpy()
    mean, var0, var1 = stats_h[:, 0], stats_h[:, 1], stats_h[:, 2]
    gmm = None
    if method == "hdt":
        pval = dip_pvalues(dip_statistics(rows, device).cpu().numpy(), n)
        bimodal = pval <= 0.05
    else:
        gmm = _gmm2(rows, device).cpu().numpy()
        # BIC = -2 * score * n + n_parameters * log(n); one component: mean, variance; two: 2 means, 2 variances, 1 weight
        v1 = var0 + 1e-6
        score1 = -0.5 * (np.log(2.0 * np.pi) + np.log(v1) + var0 / v1)
        bic1 = -2.0 * score1 * n + 2.0 * np.log(n)
        bic2 = -2.0 * gmm[:, 6] * n + 5.0 * np.log(n)
        bimodal = bic2 <= bic1
    thr = mean + 2.0 * np.sqrt(var0)
    idx = np.flatnonzero(bimodal)
    if len(idx):
        sel = rows[torch.from_numpy(idx).to(dev)].contiguous()
        g = (_gmm2(sel, device).cpu().numpy() if gmm is None else gmm[idx])
        lo = torch.from_numpy(np.ascontiguousarray(np.minimum(g[:, 2], g[:, 3]))).to(dev)
        hi = torch.from_numpy(np.ascontiguousarray(np.maximum(g[:, 2], g[:, 3]))).to(dev)
        kde_var = torch.from_numpy(np.ascontiguousarray(var1[idx] * float(n) ** (-0.4))).to(dev)
        x = torch.empty(len(idx), dtype=torch.float64, device=dev)
        _native.check(lib.aucell_binarize_trough_f64(device, sel.data_ptr(), len(idx), n, lo.data_ptr(), hi.data_ptr(),
                                                     kde_var.data_ptr(), 1e-5, 500, x.data_ptr(), _stream(dev)))
        thr[idx] = x.cpu().numpy()
    return pd.Series(index=auc_mtx.columns, data=thr)


def derive_threshold(auc_mtx: pd.DataFrame, regulon_name: str, seed=None, method: str = "hdt") -> float:
    """
    Derive threshold on the AUC values of the given regulon to binarize the cells in two clusters: "on" versus "off"
    state of the regulator.

    :param auc_mtx: The dataframe with the AUC values for all cells and regulons (n_cells x n_regulons).
    :param regulon_name: the name of the regulon for which to predict the threshold.
    :param method: The method to use to decide if the distribution of AUC values for the given regulon is not unimodel.
        Can be either Hartigan's Dip Test (HDT) or Bayesian Information Content (BIC).
    :return: The threshold on the AUC values.
    """
    assert auc_mtx is not None and not auc_mtx.empty
    assert regulon_name in auc_mtx.columns
    assert method in {"hdt", "bic"}
    return float(derive_thresholds(auc_mtx[[regulon_name]], seed=seed, method=method).iloc[0])


def binarize(auc_mtx: pd.DataFrame, threshold_overides: Optional[Mapping[str, float]] = None, seed=None,
             num_workers=1) -> (pd.DataFrame, pd.Series):
    """
    "Binarize" the supplied AUC matrix, i.e. decide if for each cells in the matrix a regulon is active or not based
    on the bimodal distribution of the AUC values for that regulon.

    :param auc_mtx: The dataframe with the AUC values for all cells and regulons (n_cells x n_regulons).
    :param threshold_overides: A dictionary that maps name of regulons to manually set thresholds.
    :param num_workers: accepted for compatibility; the regulons are processed side by side on the GPU.
    :return: A "binarized" dataframe and a series containing the AUC threshold used for each regulon.
    """
    thresholds = derive_thresholds(auc_mtx, seed=seed)
    if threshold_overides is not None:
        thresholds[list(threshold_overides.keys())] = list(threshold_overides.values())
    return (auc_mtx > thresholds).astype(int), thresholds
