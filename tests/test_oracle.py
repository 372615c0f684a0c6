# -*- coding: utf-8 -*-
"""CPU tests pinning the oracle (oracle/) against the golden vectors made by executing the reference
(tests/golden/make_golden.py), against pandas itself, and against hand-computed known answers."""
import os

import numpy as np
import pandas as pd
import pytest

from oracle import aucell_oracle as O
from pyscenic_b200.genesig import GeneSignature

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "aucell_golden.npz"))


def frame(values, prefix="g"):
    return pd.DataFrame(values, index=["c%d" % i for i in range(values.shape[0])],
                        columns=["%s%d" % (prefix, j) for j in range(values.shape[1])])


# ------------------------------------------------------------------------------------------ ranking
@pytest.mark.parametrize("seed", [0, 42, 123])
def test_rankings_adversarial_match_reference(seed):
    A = GOLD["adv_expr"]
    want = GOLD["adv_rank_seed%d" % seed]
    perm = GOLD["adv_perm_seed%d" % seed]
    # the permutation the reference drew is the legacy RandomState permutation (SURVEY 8a edge case 2)
    assert np.array_equal(perm, O.permutation_from_seed(A.shape[1], seed))
    assert np.array_equal(perm, np.random.RandomState(seed).permutation(A.shape[1]))
    got_pd = O.create_rankings(frame(A), seed=seed)
    assert got_pd.values.dtype == np.uint32
    assert np.array_equal(got_pd.values, want)
    assert list(got_pd.columns) == ["g%d" % g for g in perm]
    assert np.array_equal(O.create_rankings_np(A, perm), want)
    assert np.array_equal(O.create_rankings_c(A, perm), want)


def test_rankings_survey_known_answers():
    # SURVEY.md 8a edge case 1, verified there against pandas with the identity shuffle
    ident = np.arange(10)
    x = np.array([[4, 1, 3, np.nan, 2, 3, 0, 0, 0, np.nan]], dtype=np.float32)
    assert O.create_rankings_np(x, ident).tolist() == [[0, 4, 1, 8, 3, 2, 5, 6, 7, 9]]
    assert O.create_rankings_c(x, ident).tolist() == [[0, 4, 1, 8, 3, 2, 5, 6, 7, 9]]
    y = np.array([[0.0, -0.0, -1, np.inf, -np.inf, np.nan, 0.0, 1e-45]], dtype=np.float32)
    assert O.create_rankings_np(y, np.arange(8)).tolist() == [[2, 3, 5, 0, 6, 7, 4, 1]]
    assert O.create_rankings_c(y, np.arange(8)).tolist() == [[2, 3, 5, 0, 6, 7, 4, 1]]


def test_rankings_poisson_and_f64_match_reference():
    B = GOLD["pois_expr_u8"].astype(np.float32)
    perm = GOLD["pois_perm_seed42"]
    assert np.array_equal(O.create_rankings(frame(B), seed=42).values, GOLD["pois_rank_seed42"])
    assert np.array_equal(O.create_rankings_c(B, perm), GOLD["pois_rank_seed42"])
    C = GOLD["f64_expr"]
    permc = GOLD["f64_perm_seed5"]
    assert np.array_equal(O.create_rankings_c(C, permc), GOLD["f64_rank_seed5"])
    assert np.array_equal(O.create_rankings_np(C, permc), GOLD["f64_rank_seed5"])
    # the float64 fixture really needs float64: a float32 cast changes the ranking
    assert not np.array_equal(O.create_rankings_c(C.astype(np.float32), permc), GOLD["f64_rank_seed5"])


def test_rank_rows_are_permutations():
    # the one property the reference's own test pins (tests/test_aucell.py:31-35)
    rs = np.random.RandomState(5)
    x = rs.poisson(0.4, size=(30, 211)).astype(np.float32)
    df_rnk = O.create_rankings(frame(x))
    n = x.shape[1]
    assert len(df_rnk.sum(axis=1).unique()) == 1
    assert (df_rnk + 1).sum(axis=1).unique()[0] == (n * (n + 1)) / 2.0


def test_derive_auc_threshold_matches_reference():
    A = GOLD["adv_expr"]
    assert np.array_equal(O.derive_auc_threshold(frame(A)).values, GOLD["adv_auc_threshold_quantiles"])


# ------------------------------------------------------------------------------------------ AUC arithmetic
def test_rank_cutoff_bankers_rounding():
    # SURVEY.md 8a edge case 4
    assert O.derive_rank_cutoff(0.05, 2000) == 99
    assert O.derive_rank_cutoff(0.05, 2010) == 99      # 100.5 -> 100
    assert O.derive_rank_cutoff(0.05, 2030) == 101     # 101.5 -> 102
    assert O.derive_rank_cutoff(0.05, 20000) == 999
    assert O.derive_rank_cutoff(0.05, 25000) == 1249
    assert O.derive_rank_cutoff(0.05, 27998) == 1399
    for bad in (0.0, -0.1, 1.5):
        with pytest.raises(AssertionError):
            O.derive_rank_cutoff(bad, 1000)
    with pytest.raises(AssertionError):
        O.derive_rank_cutoff(1.0, 1000)       # cutoff 1000 > n - 1
    with pytest.raises(AssertionError):
        O.derive_rank_cutoff(0.0001, 1000)    # rounds to 0


def test_weighted_auc1d_hand_computed():
    # one cell, 10 genes, ranks of [4,1,3,NaN,2,3,0,0,0,NaN]; thr 0.4 -> cutoff round(4.0) - 1 = 3
    ranks = np.array([0, 4, 1, 8, 3, 2, 5, 6, 7, 9], dtype=np.uint32)
    cutoff = O.derive_rank_cutoff(0.4, 10)
    assert cutoff == 3
    # regulon A = {g0, g2, g5, g6}, unweighted: hits at ranks 0,1,2 -> (3-0)+(3-1)+(3-2) = 6; max = (3+1)*4
    cols = [0, 2, 5, 6]
    w = np.ones(4)
    for f in (O.weighted_auc1d, O.weighted_auc1d_c):
        assert f(ranks[cols], w, cutoff, float((cutoff + 1) * w.sum())) == 6.0 / 16.0
    # regulon B = {g2:0.5, g4:2.0, g5:0.25, g9:4.0}: hits g2 (rank 1) and g5 (rank 2):
    # x = [1,2,3], y = cumsum([0.5,0.25]) = [0.5,0.75] -> 1*0.5 + 1*0.75 = 1.25; max = 4 * 6.75 = 27
    cols = [2, 4, 5, 9]
    w = np.array([0.5, 2.0, 0.25, 4.0])
    for f in (O.weighted_auc1d, O.weighted_auc1d_c):
        assert f(ranks[cols], w, cutoff, 27.0) == 1.25 / 27.0
    # nothing below the cutoff -> 0
    assert O.weighted_auc1d_c(np.array([5, 9], dtype=np.uint32), np.ones(2), cutoff, 8.0) == 0.0
    # a gene ranked exactly at the cutoff does not count (strict <), SURVEY 8a edge case 5
    assert O.weighted_auc1d_c(np.array([3], dtype=np.uint32), np.ones(1), cutoff, 4.0) == 0.0
    # the best possible regulon scores cutoff/(cutoff+1)... times the triangle: all of ranks 0..2
    best = O.weighted_auc1d_c(np.array([0, 1, 2], dtype=np.uint32), np.ones(3), cutoff, 12.0)
    assert best == 6.0 / 12.0 and best < 1.0


def test_c_and_numpy_auc_agree_bitwise():
    rs = np.random.RandomState(9)
    for _ in range(50):
        n = rs.randint(1, 60)
        ranks = rs.permutation(500)[:n].astype(np.uint32)
        w = rs.lognormal(0, 1.5, size=n)
        cutoff = int(rs.randint(1, 400))
        m = float((cutoff + 1) * w.sum())
        assert O.weighted_auc1d(ranks, w, cutoff, m) == O.weighted_auc1d_c(ranks, w, cutoff, m)


def test_closed_form_equals_sequential_form():
    # sum(diff(x) * cumsum(w)) == sum_g w_g * (cutoff - r_g): exact for unit weights, ~1e-15 otherwise
    rs = np.random.RandomState(10)
    for _ in range(50):
        n = rs.randint(1, 200)
        ranks = rs.permutation(2000)[:n].astype(np.uint32)
        cutoff = 99
        hit = ranks < cutoff
        w1 = np.ones(n)
        m1 = float((cutoff + 1) * w1.sum())
        assert O.weighted_auc1d_c(ranks, w1, cutoff, m1) == float((cutoff - ranks[hit].astype(np.int64)).sum()) / m1
        w = rs.lognormal(0, 1, size=n)
        m = float((cutoff + 1) * w.sum())
        closed = float((w[hit] * (cutoff - ranks[hit].astype(np.int64))).sum()) / m
        assert O.weighted_auc1d_c(ranks, w, cutoff, m) == pytest.approx(closed, rel=1e-12, abs=0.0)


# ------------------------------------------------------------------------------------------ driver
def _cfg1():
    D = GOLD["cfg1_expr_u8"].astype(np.float32)
    ip, gg = GOLD["cfg1_sig_indptr"], GOLD["cfg1_sig_genes"]
    sigs = [GeneSignature(name="sig%02d" % k, gene2weight=["g%d" % g for g in gg[ip[k]:ip[k + 1]]])
            for k in range(len(ip) - 1)]
    return frame(D), sigs


def test_oracle_driver_reproduces_reference_driver_cfg1():
    df, sigs = _cfg1()
    got = O.aucell(df, sigs, auc_threshold=0.05, seed=42, num_workers=2)
    assert got.index.name == "Cell" and got.columns.name == "Regulon"
    assert list(got.columns) == [s.name for s in sigs] and list(got.index) == list(df.index)
    assert np.array_equal(got.values, GOLD["cfg1_auc"])
    got1 = O.aucell(df, sigs, auc_threshold=0.05, seed=42, num_workers=1)
    # the serial branch sorts both axes by label (SURVEY 8a edge case 10)
    assert list(got1.index) == list(GOLD["cfg1_w1_index"]) == sorted(df.index)
    assert list(got1.columns) == list(GOLD["cfg1_w1_columns"])
    assert np.array_equal(got1[got.columns].loc[got.index].values, GOLD["cfg1_auc"])
    gotn = O.aucell(df, sigs, auc_threshold=0.05, seed=42, normalize=True, num_workers=2)
    assert np.array_equal(gotn.values, GOLD["cfg1_auc_norm"])


def test_array_level_oracle_matches_frame_level():
    # auc_matrix_from_ranks (used by the GPU parity tests at sizes where DataFrames are slow) == enrichment4cells
    df, sigs = _cfg1()
    rnk = O.create_rankings(df, seed=42)
    perm = O.permutation_from_seed(df.shape[1], 42)
    inv = np.empty_like(perm); inv[perm] = np.arange(len(perm))
    ip, gg = GOLD["cfg1_sig_indptr"], GOLD["cfg1_sig_genes"]
    cols = [np.sort(inv[gg[ip[k]:ip[k + 1]]]) for k in range(len(ip) - 1)]
    w = [np.ones(len(c)) for c in cols]
    cutoff = O.derive_rank_cutoff(0.05, df.shape[1])
    got = O.auc_matrix_from_ranks(rnk.values, cols, w, [True] * len(cols), cutoff)
    assert np.array_equal(got, GOLD["cfg1_auc"])


def test_auc_value_range_and_invariances():
    df, sigs = _cfg1()
    auc = GOLD["cfg1_auc"]
    cutoff = O.derive_rank_cutoff(0.05, df.shape[1])
    assert auc.min() >= 0.0 and auc.max() < 1.0
    # gene order inside a signature does not matter
    rev = [GeneSignature(name=s.name, gene2weight=list(reversed(s.genes))) for s in sigs[:5]]
    a = O.aucell(df.iloc[:50], sigs[:5], seed=42, num_workers=2)
    b = O.aucell(df.iloc[:50], rev, seed=42, num_workers=2)
    assert np.array_equal(a.values, b.values)
    # noweights == all weights 1.0
    c = O.aucell(df.iloc[:50], sigs[:5], seed=42, noweights=True, num_workers=2)
    assert np.array_equal(a.values, c.values)
    assert cutoff == 99


def test_mismatch_gives_zero_column_and_warning(caplog):
    # reference tests/test_aucell.py:48-54
    df, sigs = _cfg1()
    fake = GeneSignature(name="test", gene2weight=list(map("FAKE{}".format, range(100))))
    with caplog.at_level("WARNING"):
        res = O.aucell(df.iloc[:20], [fake] + sigs[:3], seed=1, num_workers=1)
    assert (res["test"] == 0.0).all()
    assert "Less than 80% of the genes in test are present" in caplog.text


# ------------------------------------------------------------------------------------------ property tests
def test_property_rankings_and_aucs_hypothesis():
    from hypothesis import given, settings, strategies as st
    from hypothesis.extra import numpy as hnp

    vals = st.one_of(st.integers(0, 4).map(float), st.sampled_from([np.nan, np.inf, -np.inf, -0.0, 1e-45, -2.5]))

    @settings(max_examples=60, deadline=None)
    @given(hnp.arrays(np.float64, hnp.array_shapes(min_dims=2, max_dims=2, min_side=1, max_side=12), elements=vals),
           st.integers(0, 2**31 - 1))
    def check(x, seed):
        n = x.shape[1]
        perm = O.permutation_from_seed(n, seed)
        r_np = O.create_rankings_np(x, perm)
        r_c = O.create_rankings_c(x, perm)
        assert np.array_equal(r_np, r_c)
        # every row is a permutation of 0..n-1 (the property the reference's own test pins)
        assert np.array_equal(np.sort(r_np, axis=1), np.tile(np.arange(n, dtype=np.uint32), (x.shape[0], 1)))
        # pandas agrees (float32 to keep denormals as given)
        df = pd.DataFrame(x.astype(np.float32))
        assert np.array_equal(O.create_rankings(df, seed=seed).values, O.create_rankings_c(x.astype(np.float32), perm))
        # closed form == sequential form for unit weights, and 0 <= AUC < 1
        if n >= 4:
            cutoff = max(1, n // 2)
            for row in r_np:
                w = np.ones(n)
                m = float((cutoff + 1) * w.sum())
                a = O.weighted_auc1d_c(row, w, cutoff, m)
                hit = row < cutoff
                assert a == float((cutoff - row[hit].astype(np.int64)).sum()) / m
                assert 0.0 <= a < 1.0

    check()
