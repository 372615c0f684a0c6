#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""
Golden vectors for the cisTarget AUC + NES step (reference src/pyscenic/transform.py:175-199),
made in the build container from the reference's own TEST RESOURCES:

    /root/reference/src/resources/tests/hg19-tss-centered-10kb-10species.mc9nr.feather   (5 features x 22,284 genes, int16)
    /root/reference/src/resources/tests/c6.all.v6.1.symbols.gmt                          (189 gene sets)

    python tests/golden/make_golden_cistarget.py        # needs /root/reference

The arithmetic of ``calc_aucs`` lives in ctxcore (not installed, not vendored): the expected values are the ORACLE's
restatement of ctxcore.recovery.aucs (oracle/aucell_oracle.py), i.e. PARITY UNPINNED for that arithmetic; the inputs
(rankings, gene sets, which genes map to the database) are the reference's.  Gene symbols are replaced by their
database column numbers so that the fixture stays small.
"""
import os
import sys

import numpy as np
import pandas as pd
import pyarrow.feather as feather

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import aucell_oracle as O  # noqa: E402

RES = "/root/reference/src/resources/tests"


def main():
    t = feather.read_table(os.path.join(RES, "hg19-tss-centered-10kb-10species.mc9nr.feather")).to_pandas()
    t = t.set_index("features")
    genes = list(t.columns)
    col_of = {g: j for j, g in enumerate(genes)}
    rankings = t.to_numpy()
    assert rankings.dtype == np.int16 and rankings.shape == (5, 22284)
    total_genes = len(genes)
    thr = 0.05
    rs = np.random.RandomState(7)
    indptr, cols, wts = [0], [], []
    n_sets = 0
    with open(os.path.join(RES, "c6.all.v6.1.symbols.gmt")) as fh:
        for line in fh:
            f = line.rstrip("\n").split("\t")
            sel = sorted(col_of[g] for g in set(f[2:]) if g in col_of)      # db.load(module): the mapped genes
            if not sel:
                continue
            cols.extend(sel)
            wts.extend(np.maximum(rs.lognormal(0.0, 1.0, len(sel)), 1e-3))
            indptr.append(len(cols))
            n_sets += 1
            if n_sets == 40:
                break
    indptr = np.asarray(indptr, np.int64)
    cols = np.asarray(cols, np.int32)
    wts = np.asarray(wts, np.float64)
    auc_u = np.empty((rankings.shape[0], n_sets))
    auc_w = np.empty_like(auc_u)
    for m in range(n_sets):
        c = cols[indptr[m]:indptr[m + 1]]
        df = pd.DataFrame(rankings[:, c])
        auc_u[:, m] = O.aucs(df, total_genes, np.ones(len(c)), thr)
        auc_w[:, m] = O.aucs(df, total_genes, wts[indptr[m]:indptr[m + 1]], thr)
    nes_u = (auc_u - auc_u.mean(axis=0)) / auc_u.std(axis=0)
    np.savez_compressed(os.path.join(HERE, "cistarget_golden.npz"), rankings=rankings, total_genes=total_genes,
                        auc_threshold=thr, indptr=indptr, cols=cols, weights=wts, auc_unweighted=auc_u,
                        auc_weighted=auc_w, nes_unweighted=nes_u, features=np.asarray(t.index, dtype=str))
    print("wrote cistarget_golden.npz:", rankings.shape, n_sets, "gene sets;", "AUC range", auc_u.min(), auc_u.max())


if __name__ == "__main__":
    main()
