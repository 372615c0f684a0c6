# -*- coding: utf-8 -*-
"""Binarization (reference src/pyscenic/binarization.py:12-98).

Golden thresholds come from EXECUTING the reference with a stubbed ``diptest`` (tests/golden/make_golden_binarization.py):
they pin the mean + 2 std branch, the mixture fit, the kernel-density trough and the BIC decision.  The dip statistic is
checked against the oracle's plain-Python port of the published algorithm (AS 217); its p-value table is this
repository's own (the `diptest` package is not installed): that decision is "parity unpinned".

Bars: unimodal thresholds 1e-12 relative; bimodal thresholds 2e-5 absolute (the reference's own minimiser stops at
xatol = 1e-5, and its mixture -- which only bounds the search -- starts from a seeded k-means++ run).
"""
import os

import numpy as np
import pandas as pd
import pytest

from oracle import binarization_oracle as BO

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "binarization_golden.npz"))
TROUGH_ATOL = 2e-5


def _frame(tag):
    auc = GOLD[tag + "_auc"]
    return pd.DataFrame(auc, index=["c%d" % i for i in range(auc.shape[0])], columns=["R%d(+)" % i for i in range(auc.shape[1])])


@pytest.mark.parametrize("tag", ["a", "b"])
def test_oracle_reproduces_reference_thresholds(tag):
    auc, bimodal, want = GOLD[tag + "_auc"], GOLD[tag + "_bimodal"], GOLD[tag + "_thresholds"]
    got = np.asarray([BO.derive_threshold(auc[:, k], bool(bimodal[k]), seed=42) for k in range(auc.shape[1])])
    assert np.allclose(got, want, rtol=1e-12, atol=0)


def test_oracle_dip_known_answers():
    # equally spaced points: the empirical distribution is as close to uniform as it gets, dip = 1 / (2 n)
    for n in (4, 7, 50):
        assert abs(BO.dip_statistic(np.arange(n, dtype=float)) - 1.0 / (2 * n)) < 1e-15
    # two distant tight groups of equal size: the dip approaches its maximum 1/4
    x = np.sort(np.concatenate([np.linspace(0, 1e-3, 100), np.linspace(10, 10 + 1e-3, 100)]))
    assert 0.24 < BO.dip_statistic(x) <= 0.25
    assert BO.dip_statistic(np.ones(9)) == 1.0 / 18 and BO.dip_statistic(np.asarray([3.0])) == 0.5


def test_dip_pvalue_table():
    from pyscenic_b200.binarization import dip_pvalues

    for n in (10, 100, 1000, 5000, 100000, 2_000_000):
        d = np.linspace(0.0, 0.3, 200)
        p = dip_pvalues(d, n)
        assert p[0] == 1.0 and p[-1] == 0.0 and np.all(np.diff(p) <= 1e-15)
    # the 5 % critical value sits near 0.54 / sqrt(n) for large samples (Hartigan & Hartigan 1985, table)
    assert dip_pvalues(np.asarray([0.50 / np.sqrt(5000.0)]), 5000)[0] > 0.05 > dip_pvalues(np.asarray([0.58 / np.sqrt(5000.0)]), 5000)[0]


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b"])
def test_gpu_thresholds_match_reference_golden(tag):
    import torch

    from pyscenic_b200 import binarization as B

    df = _frame(tag)
    bimodal, want = GOLD[tag + "_bimodal"], GOLD[tag + "_thresholds"]
    rows, dev = B._sorted_rows(df, 0)
    # dip statistic: bit-level agreement with the plain-Python port is not required (different evaluation order of a
    # few products), 1e-12 is
    dips = B.dip_statistics(rows).cpu().numpy()
    want_dips = np.asarray([BO.dip_statistic(np.sort(df.iloc[:, k].values)) for k in range(df.shape[1])])
    assert np.allclose(dips, want_dips, rtol=1e-12, atol=1e-15)
    # the unambiguous columns: the dip test agrees with the branch the golden run took
    pval = B.dip_pvalues(dips, df.shape[0])
    clear = [0, 1, 2, 5, 6, 7, 8, 9]
    assert np.array_equal(pval[clear] <= 0.05, bimodal[clear])
    # thresholds, with the golden run's decisions imposed (so that every branch is compared on every column it took)
    stats = torch.empty((rows.shape[0], 4), dtype=torch.float64, device=dev)
    from pyscenic_b200 import _native
    lib = _native.load()
    _native.check(lib.aucell_binarize_stats_f64(0, rows.data_ptr(), rows.shape[0], rows.shape[1], stats.data_ptr(), None))
    torch.cuda.synchronize()
    st = stats.cpu().numpy()
    uni = st[:, 0] + 2.0 * np.sqrt(st[:, 1])
    assert np.allclose(uni[~bimodal], want[~bimodal], rtol=1e-12, atol=0)
    got = B.derive_thresholds(df, seed=42, method="hdt")
    sel = np.asarray(clear)
    assert np.allclose(got.values[sel][~bimodal[sel]], want[sel][~bimodal[sel]], rtol=1e-12, atol=0)
    assert np.abs(got.values[sel][bimodal[sel]] - want[sel][bimodal[sel]]).max() <= TROUGH_ATOL
    # BIC method: the same decisions as the reference, and thresholds within the bars
    got_bic = B.derive_thresholds(df, seed=42, method="bic")
    bic_bimodal = GOLD[tag + "_bic_bimodal"]
    is_uni = np.isclose(got_bic.values, uni, rtol=1e-12, atol=0)
    assert np.array_equal(~is_uni, bic_bimodal)
    both = bic_bimodal & bimodal
    assert np.abs(got_bic.values[both] - want[both]).max() <= TROUGH_ATOL


@pytest.mark.gpu
def test_gpu_binarize_api():
    from pyscenic_b200.binarization import binarize, derive_threshold

    df = _frame("a")
    bimodal, want_thr, want_bin = GOLD["a_bimodal"], GOLD["a_thresholds"], GOLD["a_binarized"]
    bin_mtx, thr = binarize(df, seed=42, num_workers=3)
    assert list(bin_mtx.columns) == list(df.columns) and list(bin_mtx.index) == list(df.index)
    assert list(thr.index) == list(df.columns) and bin_mtx.to_numpy().dtype.kind == "i"
    clear = [0, 1, 2, 5, 6, 7, 8, 9]
    # cells may differ only where the AUC lies within the threshold's tolerance of the threshold
    for k in clear:
        diff = bin_mtx.iloc[:, k].values != want_bin[:, k]
        assert np.all(np.abs(df.iloc[:, k].values[diff] - want_thr[k]) <= TROUGH_ATOL)
    assert np.array_equal(bin_mtx.iloc[:, 0].values, want_bin[:, 0])
    # overrides (reference argument name, typo included) replace the derived thresholds
    bin2, thr2 = binarize(df, threshold_overides={"R0(+)": 0.01, "R7(+)": 1.0})
    assert thr2["R0(+)"] == 0.01 and thr2["R7(+)"] == 1.0 and bin2["R7(+)"].sum() == 0
    assert np.array_equal(bin2["R0(+)"].values, (df["R0(+)"].values > 0.01).astype(int))
    assert abs(derive_threshold(df, "R5(+)", seed=42) - want_thr[5]) <= TROUGH_ATOL
    with pytest.raises(AssertionError):
        derive_threshold(df, "nope")
    with pytest.raises(AssertionError):
        derive_threshold(df, "R0(+)", method="kmeans")


@pytest.mark.gpu
def test_gpu_dip_random_samples_and_ties():
    import torch

    from pyscenic_b200.binarization import dip_statistics

    rs = np.random.RandomState(5)
    rows = []
    for n in (2, 3, 17, 200):
        for _ in range(6):
            kind = rs.randint(4)
            if kind == 0:
                x = rs.rand(n)
            elif kind == 1:
                x = np.round(rs.rand(n), 1)                                  # heavy ties
            elif kind == 2:
                x = rs.gamma(1.0, 1.0, n)
            else:
                x = np.resize(np.repeat(rs.rand(max(n // 4, 1)), 4), n)      # every value four times
            rows.append(np.sort(x))
    for n in (2, 3, 17, 200):
        batch = np.stack([r for r in rows if len(r) == n])
        got = dip_statistics(torch.from_numpy(batch).cuda()).cpu().numpy()
        want = np.asarray([BO.dip_statistic(r) for r in batch])
        assert np.allclose(got, want, rtol=1e-12, atol=1e-15), n
