# -*- coding: utf-8 -*-
"""cisTarget AUC + NES on the AUCell scatter kernel (reference src/pyscenic/transform.py:175-199).

Golden fixture: tests/golden/cistarget_golden.npz (made by tests/golden/make_golden_cistarget.py from the reference's
own ranking-database and GMT test resources; expected AUCs are the oracle's restatement of ctxcore.recovery.aucs)."""
import os

import numpy as np
import pandas as pd
import pytest

from oracle import aucell_oracle as O
from pyscenic_b200.genesig import GeneSignature

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cistarget_golden.npz")


def _fixture():
    z = np.load(GOLD)
    rankings = z["rankings"]
    genes = ["g%d" % j for j in range(rankings.shape[1])]
    df = pd.DataFrame(rankings, index=pd.Index(z["features"], name="features"), columns=genes)
    indptr, cols, wts = z["indptr"], z["cols"], z["weights"]
    mods_w, mods_u = [], []
    for m in range(len(indptr) - 1):
        c = cols[indptr[m]:indptr[m + 1]]
        w = wts[indptr[m]:indptr[m + 1]]
        g2w = {genes[j]: float(x) for j, x in zip(c, w)}
        g2w["not_in_database_%d" % m] = 1.0                  # db.load ignores genes the database does not hold
        mods_w.append(GeneSignature(name="M%d" % m, gene2weight=g2w))
        mods_u.append(GeneSignature(name="M%d" % m, gene2weight=list(g2w)))
    return z, df, mods_u, mods_w


def test_oracle_reproduces_cistarget_golden():
    z, df, mods_u, mods_w = _fixture()
    total, thr = int(z["total_genes"]), float(z["auc_threshold"])
    assert O.derive_rank_cutoff(thr, total) == 1113
    for m in (0, 7, 39):
        c = z["cols"][z["indptr"][m]:z["indptr"][m + 1]]
        got = O.aucs(pd.DataFrame(z["rankings"][:, c]), total, np.ones(len(c)), thr)
        np.testing.assert_array_equal(got, z["auc_unweighted"][:, m])
        got = O.aucs(pd.DataFrame(z["rankings"][:, c]), total, z["weights"][z["indptr"][m]:z["indptr"][m + 1]], thr)
        np.testing.assert_allclose(got, z["auc_weighted"][:, m], rtol=1e-12)


def test_nes_and_enriched_features_host_logic():
    from pyscenic_b200 import recovery as R

    rs = np.random.RandomState(3)
    a = pd.DataFrame(rs.rand(50, 3), index=["f%d" % i for i in range(50)], columns=["A", "B", "C"])
    a.iloc[4, 1] = 5.0
    z = R.nes(a)
    np.testing.assert_allclose(z.to_numpy(), (a.to_numpy() - a.to_numpy().mean(0)) / a.to_numpy().std(0))
    e = R.enriched_features(a, nes_threshold=3.0)
    assert list(e.index) == [("B", "f4")]
    assert e.loc[("B", "f4"), "AUC"] == 5.0 and e.loc[("B", "f4"), "NES"] >= 3.0


@pytest.mark.gpu
def test_module_aucs_matches_golden():
    from pyscenic_b200 import recovery as R

    z, df, mods_u, mods_w = _fixture()
    total, thr = int(z["total_genes"]), float(z["auc_threshold"])
    got = R.module_aucs(df, mods_u, total, thr, weighted_recovery=False)
    assert list(got.columns) == [m.name for m in mods_u] and list(got.index) == list(df.index)
    np.testing.assert_array_equal(got.to_numpy(), z["auc_unweighted"])          # unit weights: exact integers / max_auc
    np.testing.assert_allclose(R.nes(got).to_numpy(), z["nes_unweighted"], rtol=1e-9)
    got_w = R.module_aucs(df, mods_w, total, thr, weighted_recovery=True)
    np.testing.assert_allclose(got_w.to_numpy(), z["auc_weighted"], rtol=1e-6, atol=0)
    # weights are ignored unless weighted_recovery is set (transform.py:178-182)
    np.testing.assert_array_equal(R.module_aucs(df, mods_w, total, thr).to_numpy(), z["auc_unweighted"])
    # a module without any gene in the database: NaN column (the reference skips such modules)
    lone = GeneSignature(name="lone", gene2weight=["nowhere"])
    got2 = R.module_aucs(df, [mods_u[0], lone], total, thr)
    assert np.isnan(got2["lone"]).all() and np.array_equal(got2["M0"].to_numpy(), z["auc_unweighted"][:, 0])


@pytest.mark.gpu
def test_module_aucs_synthetic_database_vs_oracle():
    """A motif-sized database: 3,000 features x 12,000 genes of int32 ranks, 60 modules, rank_threshold given."""
    from pyscenic_b200 import recovery as R

    rs = np.random.RandomState(11)
    n_feat, n_genes = 3000, 12000
    ranks = np.argsort(rs.rand(n_feat, n_genes), axis=1).astype(np.int32)
    genes = ["g%d" % j for j in range(n_genes)]
    df = pd.DataFrame(ranks, index=["motif%d" % i for i in range(n_feat)], columns=genes)
    mods = []
    for m in range(60):
        sel = rs.choice(n_genes, rs.randint(5, 400), replace=False)
        mods.append(GeneSignature(name="mod%d" % m, gene2weight={genes[j]: float(w) for j, w in zip(sel, rs.rand(len(sel)) + 0.05)}))
    thr = 0.03
    got = R.module_aucs(df, mods, n_genes, thr, weighted_recovery=True, rank_threshold=5000)
    for m in (0, 13, 59):
        cols = sorted(genes.index(g) for g in mods[m].genes)
        w = np.asarray([mods[m][genes[j]] for j in cols])
        exp = O.aucs(df.iloc[:, cols], n_genes, w, thr)
        np.testing.assert_allclose(got.iloc[:, m].to_numpy(), exp, rtol=1e-6, atol=0)
    e = R.enriched_features(got, nes_threshold=3.0)
    z = R.nes(got)
    assert len(e) == int((z.to_numpy() >= 3.0).sum())
