#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""
Generates the golden vectors under tests/golden/ by EXECUTING THE REFERENCE's own code
(/root/reference/src/pyscenic/aucell.py) in the build container.  The reference cannot travel to the GPU box, so
the vectors are committed; this script documents exactly how they were made.

    python tests/golden/make_golden.py          # needs /root/reference

What is real reference code here and what is not:

* ``create_rankings`` / ``derive_auc_threshold`` / ``aucell4r`` / ``aucell`` / ``_enrichment`` are the reference's
  functions, imported unmodified from ``/root/reference/src/pyscenic/aucell.py`` and run on the installed pandas /
  numpy.  => the ranking vectors are outputs of the reference itself (PINNED).
* The reference imports three names from packages that are not installed here and not vendored in the tree:
  ``boltons.iterutils.chunked`` (a list chunker), ``ctxcore.genesig.GeneSignature`` and
  ``ctxcore.recovery.enrichment4cells``.  They are injected as stub modules: ``chunked`` is a 3-line equivalent,
  ``GeneSignature`` is this repository's record type, and ``enrichment4cells`` is the ORACLE's restatement of the
  published ctxcore 0.2.0 algorithm (oracle/aucell_oracle.py).  => the AUC vectors pin the reference's DRIVER
  (process fan-out, ordering, normalize, 80 % rule plumbing) but the per-cell AUC arithmetic inside them is the
  oracle's: PARITY UNPINNED for that arithmetic (see oracle/aucell_oracle.py).
"""
import os
import sys
import types

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import aucell_oracle as O  # noqa: E402
from pyscenic_b200.genesig import GeneSignature  # noqa: E402


def import_reference():
    bolt = types.ModuleType("boltons")
    it = types.ModuleType("boltons.iterutils")

    def chunked(src, size):
        if size <= 0:
            raise ValueError("expected a positive integer size")
        src = list(src)
        return [src[i : i + size] for i in range(0, len(src), size)]

    it.chunked = chunked
    bolt.iterutils = it
    sys.modules["boltons"] = bolt
    sys.modules["boltons.iterutils"] = it
    cx = types.ModuleType("ctxcore")
    gs = types.ModuleType("ctxcore.genesig")
    rc = types.ModuleType("ctxcore.recovery")
    gs.GeneSignature = GeneSignature
    rc.enrichment4cells = O.enrichment4cells
    sys.modules["ctxcore"] = cx
    sys.modules["ctxcore.genesig"] = gs
    sys.modules["ctxcore.recovery"] = rc
    sys.path.insert(0, "/root/reference/src")
    import pyscenic.aucell as ref

    return ref


def adversarial_matrix():
    """12 cells x 37 genes float32: ties, NaN, +-inf, -0.0, denormals, all-zero, all-equal, negative rows."""
    rs = np.random.RandomState(7)
    n = 37
    rows = []
    rows.append(np.zeros(n))                                        # all zero
    rows.append(np.full(n, 3.0))                                    # all equal
    r = rs.poisson(0.7, n).astype(float); rows.append(r)            # heavy ties
    r = rs.poisson(0.3, n).astype(float); r[5] = np.nan; r[20] = np.nan; rows.append(r)
    r = rs.randn(n); r[3] = np.inf; r[4] = -np.inf; r[9] = np.nan; rows.append(r)
    r = np.zeros(n); r[::3] = -0.0; r[1] = 1e-45; r[2] = -1e-45; r[7] = 1.0; rows.append(r)  # signed zeros, denormals
    r = -rs.poisson(1.0, n).astype(float); rows.append(r)           # non-positive
    r = rs.randn(n); rows.append(r)                                 # all distinct, mixed sign
    r = np.full(n, np.nan); rows.append(r)                          # all NaN
    r = np.full(n, np.nan); r[11] = 0.0; r[30] = 2.0; rows.append(r)
    r = np.arange(n, dtype=float); rows.append(r)                   # ascending
    r = np.arange(n, dtype=float)[::-1].copy(); rows.append(r)      # descending
    return np.asarray(rows, dtype=np.float32)


def frame(values, prefix="g"):
    return pd.DataFrame(values, index=["c%d" % i for i in range(values.shape[0])],
                        columns=["%s%d" % (prefix, j) for j in range(values.shape[1])])


def perm_of(df_rnk, df_in):
    return np.asarray([df_in.columns.get_loc(c) for c in df_rnk.columns], dtype=np.int64)


def main():
    ref = import_reference()
    out = {}

    # ---- rankings: adversarial float32, three seeds -------------------------------------------------------
    A = adversarial_matrix()
    dfA = frame(A)
    for seed in (0, 42, 123):
        r = ref.create_rankings(dfA, seed=seed)
        out["adv_rank_seed%d" % seed] = r.values.astype(np.uint32)
        out["adv_perm_seed%d" % seed] = perm_of(r, dfA)
    out["adv_expr"] = A
    out["adv_auc_threshold_quantiles"] = ref.derive_auc_threshold(dfA).values

    # ---- rankings: Poisson counts 200 x 500, seed 42 ----------------------------------------------------
    rs = np.random.RandomState(1)
    B = rs.poisson(0.5, size=(200, 500)).astype(np.float32)
    dfB = frame(B)
    r = ref.create_rankings(dfB, seed=42)
    out["pois_expr_u8"] = B.astype(np.uint8)
    out["pois_rank_seed42"] = r.values.astype(np.uint32)
    out["pois_perm_seed42"] = perm_of(r, dfB)

    # ---- rankings: float64 values that collide when cast to float32 --------------------------------------
    rs = np.random.RandomState(2)
    base = rs.poisson(1.0, size=(40, 120)).astype(np.float64)
    C = base + (rs.randint(0, 4, size=base.shape) * 1e-12) * (base > 0)   # distinct in f64, tied in f32
    dfC = frame(C)
    r = ref.create_rankings(dfC, seed=5)
    out["f64_expr"] = C
    out["f64_rank_seed5"] = r.values.astype(np.uint32)
    out["f64_perm_seed5"] = perm_of(r, dfC)

    # ---- aucell config 1 (BASELINE.json configs[0]): 1,000 x 2,000, 50 unweighted regulons, thr 0.05, seed 42
    rs = np.random.RandomState(0)
    lam = rs.uniform(0.3, 1.0, size=(1, 2000))
    D = rs.poisson(lam, size=(1000, 2000)).astype(np.float32)
    dfD = frame(D)
    sig_sizes = rs.randint(20, 201, size=50)
    sigs, members = [], []
    for k, sz in enumerate(sig_sizes):
        genes = rs.choice(2000, size=sz, replace=False)
        members.append(genes)
        sigs.append(GeneSignature(name="sig%02d" % k, gene2weight=["g%d" % g for g in genes]))
    auc_w4 = ref.aucell(dfD, sigs, auc_threshold=0.05, seed=42, num_workers=4)
    auc_w1 = ref.aucell(dfD, sigs, auc_threshold=0.05, seed=42, num_workers=1)
    assert list(auc_w4.columns) == [s.name for s in sigs]
    assert np.array_equal(auc_w1[auc_w4.columns].loc[auc_w4.index].values, auc_w4.values)
    out["cfg1_expr_u8"] = D.astype(np.uint8)
    out["cfg1_sig_indptr"] = np.concatenate([[0], np.cumsum(sig_sizes)]).astype(np.int64)
    out["cfg1_sig_genes"] = np.concatenate(members).astype(np.int32)
    out["cfg1_auc"] = auc_w4.values
    out["cfg1_w1_index"] = np.asarray(list(auc_w1.index), dtype="U8")
    out["cfg1_w1_columns"] = np.asarray(list(auc_w1.columns), dtype="U8")
    out["cfg1_auc_norm"] = ref.aucell(dfD, sigs, auc_threshold=0.05, seed=42, normalize=True, num_workers=2).values

    # ---- weighted regulons, one below the 80 % rule, one at exactly 80 %, noweights variant -----------------
    rs = np.random.RandomState(3)
    E = np.log1p(rs.negative_binomial(1, 0.6, size=(300, 800))).astype(np.float32)
    dfE = frame(E)
    wsigs, w_members, w_weights, w_names, w_extra = [], [], [], [], []
    for k in range(20):
        sz = int(rs.randint(10, 120))
        genes = rs.choice(800, size=sz, replace=False)
        w = rs.lognormal(0.0, 1.0, size=sz)
        n_fake = 0
        if k == 3:
            n_fake = sz            # 50 % overlap -> zeros + warning
        if k == 7:
            n_fake = sz // 4       # exactly 80 % when sz % 4 == 0
            if sz % 4:
                genes = genes[: sz - sz % 4]; w = w[: sz - sz % 4]; sz = len(genes); n_fake = sz // 4
        names = ["g%d" % g for g in genes] + ["FAKE%d_%d" % (k, i) for i in range(n_fake)]
        ww = np.concatenate([w, rs.lognormal(0.0, 1.0, size=n_fake)])
        nm = "TF%d(%s)" % (k, "+" if k % 2 == 0 else "-")
        wsigs.append(GeneSignature(name=nm, gene2weight=dict(zip(names, ww))))
        w_members.append(genes); w_weights.append(w); w_names.append(nm); w_extra.append(n_fake)
    auc_w = ref.aucell(dfE, wsigs, auc_threshold=0.05, seed=11, num_workers=3)
    auc_nw = ref.aucell(dfE, wsigs, auc_threshold=0.05, noweights=True, seed=11, num_workers=3)
    out["w_expr"] = E
    out["w_sig_indptr"] = np.concatenate([[0], np.cumsum([len(m) for m in w_members])]).astype(np.int64)
    out["w_sig_genes"] = np.concatenate(w_members).astype(np.int32)
    out["w_sig_weights"] = np.concatenate(w_weights).astype(np.float64)
    out["w_sig_nfake"] = np.asarray(w_extra, dtype=np.int32)
    out["w_sig_names"] = np.asarray(w_names, dtype="U16")
    out["w_auc"] = auc_w.values
    out["w_auc_noweights"] = auc_nw.values
    # aucell4r straight from a ranking frame (columns in shuffled order)
    rnkE = ref.create_rankings(dfE, seed=11)
    out["w_rank_seed11"] = rnkE.values.astype(np.uint32)
    out["w_perm_seed11"] = perm_of(rnkE, dfE)
    out["w_auc4r_thr10"] = ref.aucell4r(rnkE, wsigs, auc_threshold=0.10, num_workers=2).values

    path = os.path.join(HERE, "aucell_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB;", len(out), "arrays")


if __name__ == "__main__":
    main()
