# -*- coding: utf-8 -*-
"""
Synthetic single-cell expression matrices and regulons of the shapes BASELINE.json names.
Shared by tests/ and bench.py.  Nothing here is on the product path.

Expression values are log1p of geometric-like counts (>= 1), so a cell holds only a few dozen distinct values and
the ranking is dominated by ties -- like real scRNA-seq data, and the hard case for bit-exact tie order.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np


# ---------------------------------------------------------------------------------------------------- regulons
def make_regulons(rs: np.random.RandomState, n_genes: int, n_regulons: int, size_lo: int = 10, size_hi: int = 400,
                  weighted: bool = True, repressors: bool = True):
    """Regulon gene lists with log-uniform sizes; names alternate ``TFk(+)`` / ``TFk(-)`` (identical arithmetic,
    reference src/pyscenic/transform.py:409-415).  Returns (names, indptr int64, genes int32, weights float64)."""
    sizes = np.exp(rs.uniform(math.log(size_lo), math.log(size_hi + 1), size=n_regulons)).astype(np.int64)
    sizes = np.clip(sizes, size_lo, min(size_hi, n_genes))
    indptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    genes = np.empty(indptr[-1], dtype=np.int32)
    for k in range(n_regulons):
        genes[indptr[k]:indptr[k + 1]] = rs.choice(n_genes, size=sizes[k], replace=False)
    if weighted:
        weights = np.maximum(rs.lognormal(0.0, 1.0, size=indptr[-1]), 1e-3)
    else:
        weights = np.ones(indptr[-1], dtype=np.float64)
    names = ["TF%d(%s)" % (k, "-" if (repressors and k % 2) else "+") for k in range(n_regulons)]
    return names, indptr, genes, weights


def regulons_to_signatures(names, indptr, genes, weights, gene_names=None):
    from pyscenic_b200.genesig import GeneSignature

    sigs = []
    for k, nm in enumerate(names):
        g = genes[indptr[k]:indptr[k + 1]]
        w = weights[indptr[k]:indptr[k + 1]]
        labels = ["g%d" % x for x in g] if gene_names is None else [gene_names[x] for x in g]
        sigs.append(GeneSignature(name=nm, gene2weight=dict(zip(labels, w.tolist()))))
    return sigs


def regulon_tables_from_arrays(names, indptr, genes, weights, n_genes: int, rank_cutoff: int,
                               unweighted: bool = False):
    """RegulonTables for AUCellPlan straight from arrays (every gene is a matrix column -> all valid)."""
    from pyscenic_b200.plan import RegulonTables

    R = len(names)
    max_auc = np.empty(R, dtype=np.float64)
    for k in range(R):
        w = np.ones(indptr[k + 1] - indptr[k]) if unweighted else weights[indptr[k]:indptr[k + 1]]
        max_auc[k] = float((rank_cutoff + 1) * w.sum())
    return RegulonTables(
        names=list(names), indptr=np.asarray(indptr, dtype=np.int64), cols=np.asarray(genes, dtype=np.int32),
        weights=None if unweighted else np.asarray(weights, dtype=np.float64), max_auc=max_auc,
        valid=np.ones(R, dtype=np.uint8), rank_cutoff=int(rank_cutoff), n_genes=int(n_genes),
    )


# ---------------------------------------------------------------------------------------------------- dense
def dense_counts(rs: np.random.RandomState, n_cells: int, n_genes: int, density: float = 0.10) -> np.ndarray:
    """Dense float32 matrix, ~density non-zero, values log1p(count)."""
    nz = rs.random_sample((n_cells, n_genes)) < density
    counts = 1 + np.floor(rs.exponential(1.2, size=(n_cells, n_genes)))
    return np.where(nz, np.log1p(counts), 0.0).astype(np.float32)


# ---------------------------------------------------------------------------------------------------- CSR (torch)
def synth_csr_torch(n_cells: int, n_genes: int, mean_nnz: float, seed: int = 0, device="cpu", sigma: float = 0.35,
                    chunk_cells: int = 32768):
    """CSR expression matrix generated with torch ops on `device` (CPU for tests, CUDA for the benchmark):
    int64 indptr, int32 sorted unique column indices per row, float32 values log1p(count), count >= 1.
    Per-cell nnz is log-normal around mean_nnz, so some cells have fewer expressed genes than the rank cutoff."""
    import torch

    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    mu = math.log(mean_nnz) - 0.5 * sigma * sigma
    nn = torch.exp(torch.randn(n_cells, generator=g, device=dev, dtype=torch.float64) * sigma + mu)
    nn = nn.clamp(4, 0.6 * n_genes).to(torch.int64)
    indptr = torch.zeros(n_cells + 1, dtype=torch.int64, device=dev)
    indptr[1:] = torch.cumsum(nn, 0)
    total = int(indptr[-1].item())
    indices = torch.empty(total, dtype=torch.int32, device=dev)
    data = torch.empty(total, dtype=torch.float32, device=dev)
    for c0 in range(0, n_cells, chunk_cells):
        c1 = min(n_cells, c0 + chunk_cells)
        n = nn[c0:c1]
        e0, e1 = int(indptr[c0].item()), int(indptr[c1].item())
        tot = e1 - e0
        if tot == 0:
            continue
        start = indptr[c0:c1] - e0                                   # row starts within the chunk
        row = torch.repeat_interleave(torch.arange(c1 - c0, device=dev), n)
        k = torch.arange(tot, device=dev) - start[row]               # position within the row
        gaps = -torch.log(torch.rand(tot, generator=g, device=dev, dtype=torch.float64).clamp_min(1e-12))
        cs = torch.cumsum(gaps, 0)
        row_excl = cs[start] - gaps[start]                           # cumulative sum before each row
        within = cs - row_excl[row]                                  # inclusive cumsum inside the row
        last = cs[start + n - 1] - row_excl
        tail = -torch.log(torch.rand(c1 - c0, generator=g, device=dev, dtype=torch.float64).clamp_min(1e-12))
        frac = within / (last + tail)[row]                           # in (0, 1), increasing inside a row
        col = k + torch.floor(frac * (n_genes - n)[row].to(torch.float64)).to(torch.int64)
        indices[e0:e1] = col.clamp_(0, n_genes - 1).to(torch.int32)
        cnt = 1 + torch.floor(-torch.log(torch.rand(tot, generator=g, device=dev).clamp_min(1e-12)) * 1.2)
        data[e0:e1] = torch.log1p(cnt)
        del row, k, gaps, cs, within, frac, col, cnt
    return indptr, indices, data


def csr_to_dense(indptr: np.ndarray, indices: np.ndarray, data: np.ndarray, n_genes: int, c0: int = 0,
                 c1: Optional[int] = None) -> np.ndarray:
    n_cells = len(indptr) - 1
    c1 = n_cells if c1 is None else c1
    out = np.zeros((c1 - c0, n_genes), dtype=data.dtype)
    for c in range(c0, c1):
        s, e = indptr[c], indptr[c + 1]
        out[c - c0, indices[s:e]] = data[s:e]
    return out
