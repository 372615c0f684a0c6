# -*- coding: utf-8 -*-
"""
GPU parity tests (``-m gpu``): the CUDA path, called through the C-ABI of libaucell_b200.so, against
  (1) the committed golden vectors produced by executing the reference (tests/golden/make_golden.py),
  (2) the CPU oracle (oracle/) on seeded inputs,
  (3) size-independent properties at larger sizes.

Bars (SURVEY.md 8c / BASELINE.json north_star):
  rankings ............ bit-exact
  unweighted AUC ...... bit-exact (integer numerator, one IEEE fp64 divide)
  weighted AUC ........ relative error <= 1e-6 (RTOL below); the kernel sums w*(cutoff-rank) in a fixed tree
                        order, the reference sums diff(x)*cumsum(w) sequentially -> differences ~1e-16.
"""
import os

import numpy as np
import pandas as pd
import pytest

import synth
from oracle import aucell_oracle as O
from pyscenic_b200.genesig import GeneSignature

pytestmark = pytest.mark.gpu

RTOL = 1e-6  # tolerance stated by BASELINE.json's north_star for floating-point AUCs
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "aucell_golden.npz"))


def frame(values, prefix="g"):
    return pd.DataFrame(values, index=["c%d" % i for i in range(values.shape[0])],
                        columns=["%s%d" % (prefix, j) for j in range(values.shape[1])])


def cuda(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def gpu_rank_dense(values, perm):
    from pyscenic_b200.plan import AUCellPlan

    with AUCellPlan(values.shape[1], perm, None) as plan:
        return plan.rank_dense(cuda(values)).cpu().numpy().view(np.uint32)


def gpu_rank_csr(csr, perm):
    from pyscenic_b200.plan import AUCellPlan

    with AUCellPlan(csr.shape[1], perm, None) as plan:
        r = plan.rank_csr(cuda(csr.indptr.astype(np.int64)), cuda(csr.indices.astype(np.int32)), cuda(csr.data))
        return r.cpu().numpy().view(np.uint32)


def assert_auc_close(got, want):
    assert got.shape == want.shape
    assert np.array_equal(got == 0.0, want == 0.0)
    err = np.abs(got - want) / np.maximum(np.abs(want), 1e-300)
    err[want == 0.0] = 0.0
    assert err.max() <= RTOL, "max relative error %g" % err.max()


def oracle_auc(values, perm, reg, cutoff, unweighted=False):
    names, indptr, genes, weights = reg
    ranks = O.create_rankings_c(values, perm)
    inv = np.empty_like(perm); inv[perm] = np.arange(len(perm))
    cols, ws = [], []
    for k in range(len(names)):
        p = inv[genes[indptr[k]:indptr[k + 1]]]
        o = np.argsort(p, kind="stable")
        cols.append(p[o])
        ws.append(np.ones(len(o)) if unweighted else weights[indptr[k]:indptr[k + 1]][o])
    return O.auc_matrix_from_ranks(ranks, cols, ws, [True] * len(names), cutoff)


# =============================================================================================== rankings
@pytest.mark.parametrize("seed", [0, 42, 123])
def test_rank_dense_golden_adversarial(seed):
    A = GOLD["adv_expr"]
    got = gpu_rank_dense(A, GOLD["adv_perm_seed%d" % seed])
    assert np.array_equal(got, GOLD["adv_rank_seed%d" % seed])
    got64 = gpu_rank_dense(A.astype(np.float64), GOLD["adv_perm_seed%d" % seed])
    assert np.array_equal(got64, GOLD["adv_rank_seed%d" % seed])


def test_rank_dense_golden_poisson_and_f64():
    B = GOLD["pois_expr_u8"].astype(np.float32)
    assert np.array_equal(gpu_rank_dense(B, GOLD["pois_perm_seed42"]), GOLD["pois_rank_seed42"])
    C = GOLD["f64_expr"]
    assert np.array_equal(gpu_rank_dense(C, GOLD["f64_perm_seed5"]), GOLD["f64_rank_seed5"])


def test_create_rankings_api_matches_reference_frame():
    from pyscenic_b200.aucell import create_rankings

    A = GOLD["adv_expr"]
    df = frame(A)
    got = create_rankings(df, seed=42)
    want = O.create_rankings(df, seed=42)
    assert got.values.dtype == np.uint32
    assert list(got.columns) == list(want.columns) and list(got.index) == list(want.index)
    assert np.array_equal(got.values, want.values)
    # float64 frame whose values are not float32-representable takes the f64 kernels
    dfc = frame(GOLD["f64_expr"])
    assert np.array_equal(create_rankings(dfc, seed=5).values, GOLD["f64_rank_seed5"])
    # reference test (tests/test_aucell.py:31-35): every row is a permutation
    n = A.shape[1]
    assert (got.astype(np.int64) + 1).sum(axis=1).unique()[0] == (n * (n + 1)) / 2.0


def test_rank_csr_vs_oracle_with_explicit_zeros_negatives_nan():
    import scipy.sparse as sp

    rs = np.random.RandomState(3)
    n_cells, n_genes = 64, 700
    dense = synth.dense_counts(rs, n_cells, n_genes, density=0.15)
    dense[5, :] = 0.0                      # empty row
    dense[6, rs.choice(n_genes, 40, replace=False)] = -1.5     # negatives
    dense[7, rs.choice(n_genes, 10, replace=False)] = np.nan
    dense[8, 3] = np.inf; dense[8, 4] = -np.inf
    csr = sp.csr_matrix(dense)
    # add explicit stored zeros (and a -0.0): they must rank exactly like implicit zeros
    csr = csr.tolil(); csr = csr.tocsr()
    data = csr.data.copy(); data[::17] = 0.0; data[5::29] = -0.0
    csr = sp.csr_matrix((data, csr.indices, csr.indptr), shape=csr.shape)
    dense2 = np.asarray(csr.todense(), dtype=np.float32)
    perm = np.random.RandomState(11).permutation(n_genes)
    want = O.create_rankings_c(dense2, perm)
    assert np.array_equal(gpu_rank_csr(csr, perm), want)
    assert np.array_equal(gpu_rank_dense(dense2, perm), want)
    assert np.array_equal(gpu_rank_csr(csr.astype(np.float64), perm), want)


def test_rank_large_rows_are_permutations_and_match_oracle_sample():
    # 27,998 genes (10x mouse shape): smem list overflows for the dense-distinct rows -> global scratch path
    rs = np.random.RandomState(4)
    n_cells, n_genes = 24, 27998
    X = synth.dense_counts(rs, n_cells, n_genes, density=0.08)
    X[0] = rs.randn(n_genes).astype(np.float32)          # all distinct, mixed sign: 27,998 explicit entries
    X[1] = np.abs(rs.randn(n_genes)).astype(np.float32)  # all positive distinct
    perm = np.random.RandomState(42).permutation(n_genes)
    got = gpu_rank_dense(X, perm)
    assert np.array_equal(np.sort(got, axis=1), np.tile(np.arange(n_genes, dtype=np.uint32), (n_cells, 1)))
    assert np.array_equal(got, O.create_rankings_c(X, perm))


def test_rank_more_than_65535_genes_uses_32bit_positions():
    rs = np.random.RandomState(6)
    n_cells, n_genes = 6, 70001
    X = synth.dense_counts(rs, n_cells, n_genes, density=0.03)
    perm = np.random.RandomState(1).permutation(n_genes)
    assert np.array_equal(gpu_rank_dense(X, perm), O.create_rankings_c(X, perm))


# =============================================================================================== fused AUC
def _cfg1():
    D = GOLD["cfg1_expr_u8"].astype(np.float32)
    ip, gg = GOLD["cfg1_sig_indptr"], GOLD["cfg1_sig_genes"]
    sigs = [GeneSignature(name="sig%02d" % k, gene2weight=["g%d" % g for g in gg[ip[k]:ip[k + 1]]])
            for k in range(len(ip) - 1)]
    return frame(D), sigs


def test_aucell_cfg1_golden_bit_exact():
    """BASELINE.json configs[0]: 1,000 x 2,000, 50 unweighted regulons, auc_threshold 0.05, seed 42."""
    from pyscenic_b200.aucell import aucell

    df, sigs = _cfg1()
    got = aucell(df, sigs, auc_threshold=0.05, seed=42, num_workers=4)
    assert got.index.name == "Cell" and got.columns.name == "Regulon"
    assert list(got.index) == list(df.index) and list(got.columns) == [s.name for s in sigs]
    assert got.values.dtype == np.float64
    assert np.array_equal(got.values, GOLD["cfg1_auc"])
    # serial-branch ordering of the reference: both axes sorted by label
    got1 = aucell(df, sigs, auc_threshold=0.05, seed=42, num_workers=1)
    assert list(got1.index) == list(GOLD["cfg1_w1_index"]) and list(got1.columns) == list(GOLD["cfg1_w1_columns"])
    assert np.array_equal(got1[got.columns].loc[got.index].values, GOLD["cfg1_auc"])
    gotn = aucell(df, sigs, auc_threshold=0.05, seed=42, normalize=True)
    assert np.array_equal(gotn.values, GOLD["cfg1_auc_norm"])
    # float64 frame (what pd.read_csv yields) gives the same result
    got64 = aucell(df.astype(np.float64), sigs, auc_threshold=0.05, seed=42)
    assert np.array_equal(got64.values, GOLD["cfg1_auc"])


def _weighted_sigs():
    ip, gg, ww = GOLD["w_sig_indptr"], GOLD["w_sig_genes"], GOLD["w_sig_weights"]
    names, nfake = GOLD["w_sig_names"], GOLD["w_sig_nfake"]
    rs = np.random.RandomState(99)
    sigs = []
    for k in range(len(names)):
        g = ["g%d" % x for x in gg[ip[k]:ip[k + 1]]] + ["FAKE%d_%d" % (k, i) for i in range(nfake[k])]
        w = np.concatenate([ww[ip[k]:ip[k + 1]], rs.lognormal(0, 1, size=nfake[k])])
        sigs.append(GeneSignature(name=str(names[k]), gene2weight=dict(zip(g, w))))
    return sigs


def test_aucell_weighted_golden(caplog):
    from pyscenic_b200.aucell import aucell, aucell4r

    E = GOLD["w_expr"]
    df = frame(E)
    sigs = _weighted_sigs()
    with caplog.at_level("WARNING"):
        got = aucell(df, sigs, auc_threshold=0.05, seed=11)
    assert "Less than 80% of the genes in TF3(-) are present in the expression matrix." in caplog.text
    assert (got["TF3(-)"] == 0.0).all()              # < 80 % overlap -> zeros
    assert (got["TF7(-)"] != 0.0).any()              # exactly 80 % passes
    assert_auc_close(got.values, GOLD["w_auc"])
    got_nw = aucell(df, sigs, auc_threshold=0.05, noweights=True, seed=11)
    assert np.array_equal(got_nw.values, GOLD["w_auc_noweights"])      # integer path: bit-exact
    # aucell4r from the reference's ranking frame (columns in shuffled order), threshold 0.10
    perm = GOLD["w_perm_seed11"]
    rnk = pd.DataFrame(GOLD["w_rank_seed11"], index=df.index, columns=df.columns[perm])
    got4r = aucell4r(rnk, sigs, auc_threshold=0.10)
    assert_auc_close(got4r.values, GOLD["w_auc4r_thr10"])


def _plan_for(n_genes, perm, reg, cutoff, unweighted=False):
    from pyscenic_b200.plan import AUCellPlan

    tables = synth.regulon_tables_from_arrays(*reg, n_genes=n_genes, rank_cutoff=cutoff, unweighted=unweighted)
    return AUCellPlan(n_genes, perm, tables)


@pytest.mark.parametrize("unweighted", [False, True])
def test_fused_dense_and_csr_vs_oracle_edge_cells(unweighted, fused_variant):
    """Cells with fewer positive genes than the cutoff (zeros and even negatives fall below it), n_pos == cutoff,
    all-zero, NaN/inf cells; dense and CSR variants must agree bit-for-bit with each other."""
    import scipy.sparse as sp

    rs = np.random.RandomState(21)
    n_cells, n_genes = 96, 1500
    cutoff = O.derive_rank_cutoff(0.05, n_genes)      # 74
    X = synth.dense_counts(rs, n_cells, n_genes, density=0.12)
    X[0] = 0.0
    X[1] = 0.0; X[1, rs.choice(n_genes, 10, replace=False)] = 2.0          # n_pos = 10 < cutoff
    X[2] = 0.0; X[2, rs.choice(n_genes, cutoff, replace=False)] = rs.rand(cutoff) + 1   # n_pos == cutoff
    X[3] = 0.0; X[3, rs.choice(n_genes, cutoff + 1, replace=False)] = 1.0
    X[4] = -1.0; X[4, rs.choice(n_genes, 20, replace=False)] = 0.0        # 20 zeros on top, negatives below
    X[5] = -np.abs(rs.randn(n_genes)); X[5, :30] = 3.0                     # 30 pos, no zeros, negatives reach
    X[6, rs.choice(n_genes, 300, replace=False)] = np.nan
    X[7] = rs.randn(n_genes)                                               # fully dense, distinct
    X[8, 0] = np.inf; X[8, 1] = -np.inf
    X[9] = 5.0                                                             # all equal positive
    reg = synth.make_regulons(rs, n_genes, 40, 5, 300, weighted=True)
    perm = np.random.RandomState(7).permutation(n_genes)
    want = oracle_auc(X, perm, reg, cutoff, unweighted)
    csr = sp.csr_matrix(X)
    with _plan_for(n_genes, perm, reg, cutoff, unweighted) as plan:
        import torch

        nnz_d = torch.empty(n_cells, dtype=torch.int32, device="cuda")
        got_d = plan.aucell_dense(cuda(X), out_nnz=nnz_d).cpu().numpy()
        got_s = plan.aucell_csr(cuda(csr.indptr.astype(np.int64)), cuda(csr.indices.astype(np.int32)),
                                cuda(csr.data)).cpu().numpy()
        got_h = plan.aucell_dense_host(X, chunk_cells=17)
        got_hs = plan.aucell_csr_host(csr.indptr, csr.indices, csr.data, chunk_cells=13)
        got_d64 = plan.aucell_dense(cuda(X.astype(np.float64))).cpu().numpy()
    assert np.array_equal(got_d, got_s) and np.array_equal(got_d, got_h) and np.array_equal(got_d, got_hs)
    assert np.array_equal(got_d, got_d64)
    if unweighted:
        assert np.array_equal(got_d, want)
    else:
        assert_auc_close(got_d, want)
    # np.count_nonzero semantics (derive_auc_threshold): NaN counts, -0.0 does not
    assert np.array_equal(nnz_d.cpu().numpy(), np.count_nonzero(X, axis=1))


@pytest.fixture(params=["default"])
def fused_variant(request):
    """Hook for A/B builds of kernel variants; the tree holds one fused kernel."""
    return request.param


@pytest.mark.parametrize("unweighted", [False, True])
def test_fast_path_and_its_fallback(unweighted, fused_variant):
    """The bucket-sort fast path and the cells it must hand to the general sort: many distinct near-equal values
    sharing one value bin (crowded bucket), binary data (one tie group), all-distinct data, a +inf outlier."""
    import scipy.sparse as sp

    rs = np.random.RandomState(77)
    n_cells, n_genes = 64, 6000
    cutoff = O.derive_rank_cutoff(0.05, n_genes)      # 299
    X = synth.dense_counts(rs, n_cells, n_genes, density=0.2)
    idx = rs.choice(n_genes, 900, replace=False)
    X[0] = 0.0; X[0, idx] = (1.0 + np.arange(900) * 1e-6).astype(np.float32); X[0, idx[0]] = 1e6   # crowded bin
    X[1] = (rs.rand(n_genes) < 0.3).astype(np.float32)                       # binary: one tie group
    X[2] = 0.0; X[2, idx] = rs.rand(900).astype(np.float32) + 0.5             # all distinct
    X[3, 5] = np.inf
    X[4] = 0.0; X[4, idx[:400]] = 2.0; X[4, idx[400:]] = np.nextafter(np.float32(2.0), np.float32(3.0))  # 2 keys, 1 bin
    X[5] = 0.0; X[5, idx[:50]] = np.float32(1e-45)                            # denormals only, n_pos < cutoff
    X[6] = np.abs(rs.randn(n_genes)).astype(np.float32) + 1.0                 # dense positive: select mode
    X[7] = -np.abs(rs.randn(n_genes)).astype(np.float32)                      # select mode + zeros below the cutoff:
    X[7, idx[:100]] = rs.rand(100).astype(np.float32) + 1; X[7, idx[100:]] = 0.0   # 100 pos, 800 zeros, 5100 neg
    X[8] = rs.poisson(3.0, n_genes).astype(np.float32)                        # 95 % non-zero counts: long row, ties
    X[9] = 7.0                                                                # one value everywhere
    # two adjacent float32 values, 450 genes each, sharing one value bin with nothing to tell them apart at the bin's
    # resolution: a crowded sub-bucket whatever the split -> general sort
    X[10] = 0.0; X[10, idx[:450]] = 1.0; X[10, idx[450:]] = np.nextafter(np.float32(1.0), np.float32(2.0)); X[10, idx[0]] = 1e6
    X[11] = rs.randn(n_genes).astype(np.float32)                              # z-scored: all distinct, half negative
    X[12] = rs.rand(n_genes).astype(np.float32) * (rs.rand(n_genes) < 0.3)     # uniform (0,1): many keys per value bin
    reg = synth.make_regulons(rs, n_genes, 50, 10, 400, weighted=True)
    perm = np.random.RandomState(5).permutation(n_genes)
    want = oracle_auc(X, perm, reg, cutoff, unweighted)
    csr = sp.csr_matrix(X)
    with _plan_for(n_genes, perm, reg, cutoff, unweighted) as plan:
        got_d = plan.aucell_dense(cuda(X)).cpu().numpy()
        got_s = plan.aucell_csr(cuda(csr.indptr.astype(np.int64)), cuda(csr.indices.astype(np.int32)),
                                cuda(csr.data)).cpu().numpy()
        got_64 = plan.aucell_dense(cuda(X.astype(np.float64))).cpu().numpy()
    assert np.array_equal(got_d, got_s) and np.array_equal(got_d, got_64)
    if unweighted:
        assert np.array_equal(got_d, want)
    else:
        assert_auc_close(got_d, want)


def test_fused_csr_config3_shape_vs_oracle(fused_variant):
    """BASELINE.json configs[2] shape at oracle-friendly scale: 25k genes CSR ~7 % nnz, 500 regulons incl.
    repressor (-) regulons, weighted; per-cell nnz log-normal so some cells sit below the cutoff of 1,249."""
    n_cells, n_genes = 256, 25000
    indptr, indices, data = synth.synth_csr_torch(n_cells, n_genes, 1750, seed=5, device="cuda", sigma=0.6)
    rs = np.random.RandomState(8)
    reg = synth.make_regulons(rs, n_genes, 500, 10, 400, weighted=True, repressors=True)
    cutoff = O.derive_rank_cutoff(0.05, n_genes)
    assert cutoff == 1249
    perm = np.random.RandomState(42).permutation(n_genes)
    ip, ix, dv = indptr.cpu().numpy(), indices.cpu().numpy(), data.cpu().numpy()
    assert (np.diff(ip) < cutoff).any() and (np.diff(ip) > cutoff).any()
    X = synth.csr_to_dense(ip, ix, dv, n_genes)
    want = oracle_auc(X, perm, reg, cutoff)
    with _plan_for(n_genes, perm, reg, cutoff) as plan:
        got = plan.aucell_csr(indptr, indices, data).cpu().numpy()
        got2 = plan.aucell_csr(indptr, indices, data).cpu().numpy()
    assert_auc_close(got, want)
    assert np.array_equal(got, got2)      # deterministic
    want_u = oracle_auc(X, perm, reg, cutoff, unweighted=True)
    with _plan_for(n_genes, perm, reg, cutoff, unweighted=True) as plan:
        got_u = plan.aucell_csr(indptr, indices, data).cpu().numpy()
    assert np.array_equal(got_u, want_u)


def test_fused_dense_config2_shape_vs_oracle():
    """BASELINE.json configs[1] shape at oracle-friendly scale: 20k genes dense f32, 300 weighted regulons."""
    rs = np.random.RandomState(12)
    n_cells, n_genes = 128, 20000
    X = synth.dense_counts(rs, n_cells, n_genes, density=0.10)
    reg = synth.make_regulons(rs, n_genes, 300, 10, 1000, weighted=True)
    cutoff = O.derive_rank_cutoff(0.05, n_genes)
    assert cutoff == 999
    perm = np.random.RandomState(42).permutation(n_genes)
    want = oracle_auc(X, perm, reg, cutoff)
    with _plan_for(n_genes, perm, reg, cutoff) as plan:
        got = plan.aucell_dense(cuda(X)).cpu().numpy()
    assert_auc_close(got, want)


@pytest.mark.parametrize("thr", [0.01, 0.2, 0.6])
def test_threshold_sweep_vs_oracle(thr):
    """BASELINE.json configs[4]: cutoff-bound and gather-bound regimes (incl. a cutoff far above nnz)."""
    rs = np.random.RandomState(31)
    n_cells, n_genes = 48, 5000
    X = synth.dense_counts(rs, n_cells, n_genes, density=0.07)
    reg = synth.make_regulons(rs, n_genes, 60, 10, 600, weighted=True)
    cutoff = O.derive_rank_cutoff(thr, n_genes)
    perm = np.random.RandomState(3).permutation(n_genes)
    want = oracle_auc(X, perm, reg, cutoff)
    with _plan_for(n_genes, perm, reg, cutoff) as plan:
        got = plan.aucell_dense(cuda(X)).cpu().numpy()
    assert_auc_close(got, want)


def test_sharding_invariance_and_value_range():
    """Output is independent of how cells are split (what the multi-GPU driver relies on), 0 <= AUC < 1."""
    n_cells, n_genes = 2000, 27998
    indptr, indices, data = synth.synth_csr_torch(n_cells, n_genes, 1960, seed=9, device="cuda")
    rs = np.random.RandomState(2)
    reg = synth.make_regulons(rs, n_genes, 500, 10, 400, weighted=True)
    cutoff = O.derive_rank_cutoff(0.05, n_genes)
    perm = np.random.RandomState(0).permutation(n_genes)
    with _plan_for(n_genes, perm, reg, cutoff) as plan:
        whole = plan.aucell_csr(indptr, indices, data).cpu().numpy()
        parts = []
        for (c0, c1) in [(0, 700), (700, 701), (701, 2000)]:
            e0 = int(indptr[c0])
            ip = (indptr[c0:c1 + 1] - e0).contiguous()
            e1 = int(indptr[c1])
            parts.append(plan.aucell_csr(ip, indices[e0:e1].contiguous(), data[e0:e1].contiguous()).cpu().numpy())
    assert np.array_equal(np.concatenate(parts), whole)
    assert whole.min() >= 0.0 and whole.max() < 1.0 and np.isfinite(whole).all()


def test_sparse_api_equals_dense_api_and_mismatch_behaviour(caplog):
    import scipy.sparse as sp

    from pyscenic_b200.aucell import aucell

    df, sigs = _cfg1()
    df = df.iloc[:100]
    fake = GeneSignature(name="test", gene2weight=list(map("FAKE{}".format, range(100))))
    with caplog.at_level("WARNING"):
        d = aucell(df, [fake] + sigs[:10], seed=3)           # reference tests/test_aucell.py:48-54
    assert "Less than 80% of the genes in test are present" in caplog.text
    assert (d["test"] == 0.0).all()
    s = aucell(sp.csr_matrix(df.values), [fake] + sigs[:10], seed=3, genes=list(df.columns), cells=list(df.index))
    assert np.array_equal(d.values, s.values) and list(d.index) == list(s.index)
    want = O.aucell(df, [fake] + sigs[:10], seed=3, num_workers=2)
    assert np.array_equal(d.values, want.values)
    with pytest.raises(AssertionError):
        aucell(df, sigs[:2], auc_threshold=1.5, seed=3)
    with pytest.raises(ValueError):
        aucell(df, [], seed=3)


def test_enrichment_alias_matches_oracle():
    from pyscenic_b200.aucell import enrichment

    df, sigs = _cfg1()
    rnk = O.create_rankings(df.iloc[:64], seed=5)
    got = enrichment(rnk, sigs[0], auc_threshold=0.05)
    want = O.enrichment4cells(rnk, sigs[0], auc_threshold=0.05)
    assert list(got.index.names) == ["Cell", "Regulon"] and list(got.columns) == ["AUC"]
    assert got.index.equals(want.index)
    assert np.array_equal(got["AUC"].values, want["AUC"].values)


def test_config3_full_size_properties_and_spot_check():
    """BASELINE.json configs[2] at FULL size (100k cells x 25k genes CSR, 500 regulons): size-independent properties
    (determinism, sharding invariance, value range, permutation-invariance of the cell order) plus an oracle
    spot-check of 192 cells drawn from the whole matrix."""
    import torch

    n_cells, n_genes = 100_000, 25_000
    indptr, indices, data = synth.synth_csr_torch(n_cells, n_genes, 1750, seed=21, device="cuda")
    rs = np.random.RandomState(5)
    reg = synth.make_regulons(rs, n_genes, 500, 10, 400, weighted=True, repressors=True)
    cutoff = O.derive_rank_cutoff(0.05, n_genes)
    perm = np.random.RandomState(42).permutation(n_genes)
    with _plan_for(n_genes, perm, reg, cutoff) as plan:
        a = plan.aucell_csr(indptr, indices, data)
        b = plan.aucell_csr(indptr, indices, data)
        assert torch.equal(a, b)                                        # deterministic
        stats, fast = plan.path_stats(), plan.fast_stats()
        assert stats["cells"] + fast["ranked"] == 2 * n_cells           # every cell ranked once, by one of the two kernels
        assert fast["ranked"] > 1.9 * n_cells                             # count data: the tie-class kernel takes it
        # two halves computed separately == the whole (what cell sharding over GPUs relies on)
        mid = 41_337
        e_mid = int(indptr[mid])
        lo = plan.aucell_csr(indptr[: mid + 1].contiguous(), indices[:e_mid].contiguous(), data[:e_mid].contiguous())
        hi = plan.aucell_csr((indptr[mid:] - e_mid).contiguous(), indices[e_mid:].contiguous(), data[e_mid:].contiguous())
        assert torch.equal(torch.cat([lo, hi]), a)
        # the host-buffer entry point gives the same matrix
        h = plan.aucell_csr_host(indptr.cpu().numpy(), indices.cpu().numpy(), data.cpu().numpy())
        assert np.array_equal(h, a.cpu().numpy())
    av = a.cpu().numpy()
    assert np.isfinite(av).all() and av.min() >= 0.0 and av.max() < 1.0
    pick = np.random.RandomState(0).choice(n_cells, 192, replace=False)
    ip = indptr.cpu().numpy()
    ix, dv = indices.cpu().numpy(), data.cpu().numpy()
    X = np.zeros((len(pick), n_genes), dtype=np.float32)
    for r, c in enumerate(pick):
        X[r, ix[ip[c]:ip[c + 1]]] = dv[ip[c]:ip[c + 1]]
    assert_auc_close(av[pick], oracle_auc(X, perm, reg, cutoff))


def _full_size_check(n_cells, n_genes, mean_nnz, n_regulons, thr, seed, n_sample=2048):
    """Shared body of the full-size parity tests: run-to-run bit identity, the tie-class kernel against the general
    kernel on EVERY cell, value range, and a random sample of cells against the CPU oracle."""
    import torch

    indptr, indices, data = synth.synth_csr_torch(n_cells, n_genes, mean_nnz, seed=seed, device="cuda")
    rs = np.random.RandomState(seed)
    reg = synth.make_regulons(rs, n_genes, n_regulons, 10, 400, weighted=True, repressors=True)
    cutoff = O.derive_rank_cutoff(thr, n_genes)
    perm = np.random.RandomState(42).permutation(n_genes)
    with _plan_for(n_genes, perm, reg, cutoff) as plan:
        a = plan.aucell_csr(indptr, indices, data)
        fast = plan.fast_stats()
        assert torch.equal(plan.aucell_csr(indptr, indices, data), a)          # deterministic
        plan.set_fast_path(False)
        assert torch.equal(plan.aucell_csr(indptr, indices, data), a)          # both kernels, every cell, every bit
        plan.set_fast_path(True)
    assert fast["ranked"] > 0.9 * n_cells
    assert bool(torch.isfinite(a).all()) and float(a.min()) >= 0.0 and float(a.max()) < 1.0
    pick = np.sort(np.random.RandomState(seed + 1).choice(n_cells, n_sample, replace=False))
    ip = indptr.cpu().numpy()
    ix, dv = indices.cpu().numpy(), data.cpu().numpy()
    X = np.zeros((len(pick), n_genes), dtype=np.float32)
    for r, c in enumerate(pick):
        X[r, ix[ip[c]:ip[c + 1]]] = dv[ip[c]:ip[c + 1]]
    assert_auc_close(a[torch.from_numpy(pick).cuda()].cpu().numpy(), oracle_auc(X, perm, reg, cutoff))


def test_config4_shape_200k_cells_full_parity():
    """BASELINE.json configs[3] (the benchmark's shape: 27,998 genes, ~1,960 stored entries per cell, 500 weighted
    regulons incl. (-), threshold 0.05) on 200 k cells."""
    _full_size_check(200_000, 27_998, 1960, 500, 0.05, seed=404)


@pytest.mark.parametrize("thr,n_regulons", [(0.01, 50), (0.01, 2000), (0.2, 50), (0.2, 2000)])
def test_config5_corners_200k_cells_full_parity(thr, n_regulons):
    """BASELINE.json configs[4]: the corners of the threshold (0.01 .. 0.2) x regulon-count (50 .. 2,000) sweep on
    200 k cells x 20 k genes."""
    _full_size_check(200_000, 20_000, 1400, n_regulons, thr, seed=500 + n_regulons + int(thr * 100))


def test_derive_auc_threshold_dense_golden_and_sparse():
    """pyscenic_b200.aucell.derive_auc_threshold itself (not the oracle's copy): the reference's quantiles on the golden
    adversarial matrix (NaN counts as non-zero, -0.0 does not), and the same numbers from a scipy CSR matrix whose
    stored entries include explicit zeros (counted on the GPU, aucell_count_nonzero_csr_f32)."""
    import scipy.sparse as sp

    from pyscenic_b200.aucell import derive_auc_threshold

    A = GOLD["adv_expr"]
    got = derive_auc_threshold(frame(A))
    assert np.array_equal(got.values, GOLD["adv_auc_threshold_quantiles"])
    assert list(got.index) == [0.01, 0.05, 0.10, 0.50, 1]
    m = sp.csr_matrix(np.nan_to_num(A, nan=7.0))         # (scipy drops nothing: NaN would be kept as a stored entry too)
    m.data[::5] = 0.0                                    # explicit zeros among the stored entries
    m.data[3::11] = np.nan
    dense = m.toarray()
    want = pd.Series(np.count_nonzero(dense, axis=1)).quantile([0.01, 0.05, 0.10, 0.50, 1]) / dense.shape[1]
    got_s = derive_auc_threshold(m)
    assert np.array_equal(got_s.values, want.values)
    rs = np.random.RandomState(1)
    big = sp.random(3000, 5000, density=0.07, format="csr", random_state=rs, dtype=np.float32)
    want_b = pd.Series(np.diff(big.indptr)).quantile([0.01, 0.05, 0.10, 0.50, 1]) / 5000
    assert np.array_equal(derive_auc_threshold(big).values, want_b.values)
    assert np.array_equal(derive_auc_threshold(big.astype(np.float64)).values, want_b.values)


@pytest.mark.parametrize("n_genes", [60000, 70001])
def test_deep_cells_whose_selection_exceeds_the_register_capacity(n_genes):
    """Cells with more stored entries than the register capacity (4,096) whose tie class at the cutoff is so large
    that even the selection exceeds it (they take the general sort today; pruning the straddling class by shuffled
    position is the planned fix), 16- and 32-bit positions.  One row's straddling bin holds two adjacent keys."""
    import torch

    n_cells = 48
    indptr, indices, data = synth.synth_csr_torch(n_cells, n_genes, 6000, seed=41, device="cuda", sigma=0.25)
    ip = indptr.cpu().numpy()
    assert (np.diff(ip) > 4500).sum() >= n_cells // 2 and np.diff(ip)[3] > 4500
    dv = data.cpu().numpy().copy()
    # cell 3: the lowest class gets a second, adjacent value -> the straddling bin holds two keys
    s3, e3 = ip[3], ip[4]
    lo = dv[s3:e3].min()
    sel = np.flatnonzero(dv[s3:e3] == lo)[::2]
    dv[s3 + sel] = np.nextafter(np.float32(lo), np.float32(10.0))
    data = torch.from_numpy(dv).cuda()
    rs = np.random.RandomState(6)
    reg = synth.make_regulons(rs, n_genes, 80, 10, 400, weighted=True)
    cutoff = O.derive_rank_cutoff(0.05, n_genes)
    perm = np.random.RandomState(42).permutation(n_genes)
    X = synth.csr_to_dense(ip, indices.cpu().numpy(), dv, n_genes)
    want = oracle_auc(X, perm, reg, cutoff)
    with _plan_for(n_genes, perm, reg, cutoff) as plan:
        got = plan.aucell_csr(indptr, indices, data).cpu().numpy()
        stats, fast = plan.path_stats(), plan.fast_stats()
        got_d = plan.aucell_dense(torch.from_numpy(X).cuda()).cpu().numpy()
    assert_auc_close(got, want)
    assert np.array_equal(got, got_d)
    assert stats["cells"] + fast["ranked"] == n_cells


def _mixed_bag_matrix(rs, n_genes):
    """CSR rows that exercise every exit of the tie-class kernel: count-like rows (its home ground), an empty row, a row
    with fewer entries than the cutoff (zero fill), explicit zeros, negatives, NaN, +inf, a continuous row, a row with
    more distinct values than classes (tail), more than 128 tail values, two values in one bin, a row longer than the
    kernel's capacity, a one-entry row."""
    import scipy.sparse as sp

    rows = []

    def counts(n, lam=1.2):
        v = np.zeros(n_genes, dtype=np.float32)
        idx = rs.choice(n_genes, n, replace=False)
        v[idx] = np.log1p(1 + np.floor(rs.exponential(lam, size=n))).astype(np.float32)
        return v

    for n in (1200, 300, 40, 2500, 1):
        rows.append(counts(min(n, n_genes // 2)))
    rows.append(np.zeros(n_genes, dtype=np.float32))                                     # empty
    r = counts(800); r[rs.choice(n_genes, 30, replace=False)] = -2.5; rows.append(r)     # negatives
    r = counts(800); r[rs.choice(n_genes, 5, replace=False)] = np.nan; rows.append(r)    # NaN
    r = counts(800); r[rs.choice(n_genes, 3, replace=False)] = np.inf; rows.append(r)    # +inf
    r = np.zeros(n_genes, dtype=np.float32); idx = rs.choice(n_genes, 900, replace=False)
    r[idx] = rs.lognormal(0, 1, 900).astype(np.float32); rows.append(r)                  # continuous
    r = np.zeros(n_genes, dtype=np.float32); idx = rs.choice(n_genes, 1500, replace=False)
    r[idx] = (1 + np.floor(rs.exponential(12.0, 1500))).astype(np.float32); rows.append(r)   # ~60 distinct raw counts
    r = np.zeros(n_genes, dtype=np.float32); idx = rs.choice(n_genes, 1500, replace=False)
    r[idx] = (1 + np.floor(rs.exponential(200.0, 1500))).astype(np.float32); rows.append(r)  # hundreds of distinct values
    r = counts(900); idx = np.flatnonzero(r)[:200]
    r[idx[:100]] = np.float32(1.0); r[idx[100:]] = np.nextafter(np.float32(1.0), np.float32(2.0)); rows.append(r)  # one bin, two values
    rows.append(counts(min(4600, n_genes - 10)))                                          # longer than the capacity
    X = np.stack(rows)
    m = sp.csr_matrix(X)
    # explicit zeros in a count-like row (stored entries whose value is 0.0 / -0.0)
    z = sp.csr_matrix(counts(700)[None, :])
    z.data[::7] = 0.0
    z.data[3::50] = -0.0
    m = sp.vstack([m, z]).tocsr()
    X = np.vstack([X, z.toarray()])
    return X, m


@pytest.mark.parametrize("tie_kernel", ["12", "1", "2"])
@pytest.mark.parametrize("unweighted", [False, True])
@pytest.mark.parametrize("n_genes", [5000, 20011])
def test_tie_class_kernel_equals_general_kernel_bit_for_bit(unweighted, n_genes, tie_kernel, monkeypatch):
    """The kernels behind aucell_csr_f32 / aucell_csr16_f32 add the same integers: switching the tie-class kernels
    off must not change one bit, whatever mixture of cells the matrix holds; and all match the oracle.
    tie_kernel: "1" aucell_tie_kernel alone, "2" aucell_tie2_kernel alone (capacity lowered so that it hands the long
    row over), "12" the default -- the first kernel with the rows beyond its capacity handed to the second."""
    import torch

    monkeypatch.setenv("AUCELL_B200_TIE_KERNEL", tie_kernel)
    if tie_kernel == "2":
        monkeypatch.setenv("AUCELL_B200_TIE2_CAP", "4096")

    rs = np.random.RandomState(11)
    X, m = _mixed_bag_matrix(rs, n_genes)
    reps = 40                                            # enough cells for every cell group of every CTA
    order = rs.permutation(np.tile(np.arange(X.shape[0]), reps))
    m = m[order]
    reg = synth.make_regulons(rs, n_genes, 120, 10, 400, weighted=not unweighted)
    cutoff = O.derive_rank_cutoff(0.05, n_genes)
    perm = np.random.RandomState(3).permutation(n_genes)
    indptr, indices, data = cuda(m.indptr.astype(np.int64)), cuda(m.indices.astype(np.int32)), cuda(m.data)
    with _plan_for(n_genes, perm, reg, cutoff, unweighted=unweighted) as plan:
        assert plan.set_fast_path(True)
        nnz_f = torch.empty(m.shape[0], dtype=torch.int32, device="cuda")
        fast = plan.aucell_csr(indptr, indices, data, out_nnz=nnz_f)
        fs = plan.fast_stats()
        fast16 = plan.aucell_csr(indptr, indices.to(torch.uint16), data)
        plan.set_fast_path(False)
        nnz_g = torch.empty_like(nnz_f)
        general = plan.aucell_csr(indptr, indices, data, out_nnz=nnz_g)
        plan.set_fast_path(True)
    assert fs["ranked"] > 0 and fs["handed_negative_or_nan"] > 0
    assert (fs["handed_long_row"] > 0) == (tie_kernel != "12")
    assert fs["handed_mixed_bin"] + fs["handed_tail_overflow"] + fs["handed_skipped"] > 0
    assert fs["ranked"] + sum(v for k, v in fs.items() if k.startswith("handed")) == m.shape[0]
    assert torch.equal(fast, general) and torch.equal(fast16, general)
    assert torch.equal(nnz_f, nnz_g)
    want = oracle_auc(X, perm, reg, cutoff, unweighted=unweighted)[order]
    if unweighted:
        assert np.array_equal(fast.cpu().numpy(), want)
    else:
        assert_auc_close(fast.cpu().numpy(), want)


@pytest.mark.parametrize("unweighted", [False, True])
def test_byte_coded_values_device_and_host_paths(unweighted, monkeypatch):
    """3 bytes per stored entry: uint16 column index + one-byte code into a table of the distinct values
    (aucell_csr16_lut8_f32, what the host pipeline uploads while the matrix holds <= 256 distinct values).  Bit-identical
    to the float32 entry point -- on the mixed bag of cells too (negatives, NaN, long rows, ... go to the general kernel
    through the decoded values) --, for dictionaries that grow from chunk to chunk, and when a later chunk overflows
    the dictionary (the call then continues with float32 values)."""
    import torch

    n_genes = 9001
    rs = np.random.RandomState(23)
    X, m = _mixed_bag_matrix(rs, n_genes)
    # the continuous row and the hundreds-of-distinct-values row do not fit one byte: drop them for the device-level call
    keep = [i for i in range(m.shape[0]) if len(np.unique(m[i].data)) <= 60]
    order = rs.permutation(np.tile(np.asarray(keep), 25))
    mk = m[order]
    vals, codes = np.unique(mk.data.view(np.uint32), return_inverse=True)
    assert len(vals) <= 256
    reg = synth.make_regulons(rs, n_genes, 90, 10, 300, weighted=not unweighted)
    cutoff = O.derive_rank_cutoff(0.05, n_genes)
    perm = np.random.RandomState(5).permutation(n_genes)
    indptr, indices, data = cuda(mk.indptr.astype(np.int64)), cuda(mk.indices.astype(np.int32)), cuda(mk.data)
    with _plan_for(n_genes, perm, reg, cutoff, unweighted=unweighted) as plan:
        want = plan.aucell_csr(indptr, indices, data)
        shuffled = rs.permutation(len(vals))                     # codes need not be ordered
        inv = np.empty_like(shuffled); inv[shuffled] = np.arange(len(vals))
        got = plan.aucell_csr_lut8(indptr, indices.to(torch.uint16), cuda(inv[codes].astype(np.uint8)),
                                   cuda(vals[shuffled].view(np.float32)))
        assert torch.equal(got, want)
        plan.set_fast_path(False)                                # every cell through the decoded values
        assert torch.equal(plan.aucell_csr_lut8(indptr, indices.to(torch.uint16), cuda(inv[codes].astype(np.uint8)),
                                                cuda(vals[shuffled].view(np.float32))), want)
        plan.set_fast_path(True)
        # host pipeline: the whole bag (the continuous row overflows the dictionary in whatever chunk it sits)
        order2 = rs.permutation(np.tile(np.arange(m.shape[0]), 25))
        m2 = m[order2]
        ip2, ix2, dv2 = m2.indptr.astype(np.int64), m2.indices.astype(np.int32), m2.data.astype(np.float32)
        want_nnz2 = torch.empty(m2.shape[0], dtype=torch.int32, device="cuda")
        want2 = plan.aucell_csr(cuda(ip2), cuda(ix2), cuda(dv2), out_nnz=want_nnz2).cpu().numpy()
        want_nnz2 = want_nnz2.cpu().numpy()
        results = {}
        for mode in ("0", "1", None):
            if mode is None:
                monkeypatch.delenv("AUCELL_B200_HOST_VAL8", raising=False)
            else:
                monkeypatch.setenv("AUCELL_B200_HOST_VAL8", mode)
            for chunk in (0, 37):
                got2, nnz2 = plan.aucell_csr_host(ip2, ix2, dv2, want_nnz=True, chunk_cells=chunk)
                assert np.array_equal(got2, want2), (mode, chunk)
                assert np.array_equal(nnz2, want_nnz2), (mode, chunk)
                results[(mode, chunk)] = plan.last_host_bytes()[0]
        # count-like rows only: the byte codes really are what goes up (3 bytes per entry + the tables)
        mk_ip, mk_ix, mk_dv = mk.indptr.astype(np.int64), mk.indices.astype(np.int32), mk.data.astype(np.float32)
        monkeypatch.setenv("AUCELL_B200_HOST_VAL8", "1")
        got3 = plan.aucell_csr_host(mk_ip, mk_ix, mk_dv, chunk_cells=64)
        sent = plan.last_host_bytes()[0]
        assert np.array_equal(got3, want.cpu().numpy())
        n_chunks = -(-mk.shape[0] // 64)
        assert sent == 8 * (mk.shape[0] + n_chunks) + 3 * mk.nnz + 1024 * n_chunks
    monkeypatch.delenv("AUCELL_B200_HOST_VAL8", raising=False)


@pytest.mark.parametrize("thr,unweighted", [(0.05, False), (0.6, False), (0.6, True)])
def test_from_rankings_hit_list_and_its_overflow(thr, unweighted):
    """aucell4r kernel: ranks below the cutoff are collected into a 4,096-entry list before they are scattered; with
    thr 0.6 of 12,001 genes a row holds 7,200 of them, so the overflow branch (scatter where found) runs too.  Rows at
    every 16-byte misalignment (odd n_genes, row_stride > n_genes)."""
    import torch

    rs = np.random.RandomState(23)
    n_cells, n_genes = 37, 12001
    X = synth.dense_counts(rs, n_cells, n_genes, density=0.15)
    perm = rs.permutation(n_genes)
    ranks = O.create_rankings_c(X, perm)                       # [n_cells, n_genes] uint32, shuffled column order
    cutoff = O.derive_rank_cutoff(thr, n_genes)
    assert (thr > 0.5) == (cutoff > 4096)
    names, indptr, genes, weights = synth.make_regulons(rs, n_genes, 90, 5, 600, weighted=True)
    cols = [np.sort(genes[indptr[k]:indptr[k + 1]]) for k in range(len(names))]
    ws = [np.ones(len(c)) if unweighted else weights[indptr[k]:indptr[k + 1]][np.argsort(genes[indptr[k]:indptr[k + 1]], kind="stable")]
          for k, c in enumerate(cols)]
    want = O.auc_matrix_from_ranks(ranks, cols, ws, [True] * len(names), cutoff)
    tables = synth.regulon_tables_from_arrays(names, indptr, genes, weights, n_genes=n_genes, rank_cutoff=cutoff,
                                              unweighted=unweighted)
    from pyscenic_b200.plan import AUCellPlan

    with AUCellPlan(n_genes, None, tables) as plan:            # identity permutation: columns are ranking columns
        padded = torch.zeros((n_cells, n_genes + 3), dtype=torch.int32, device="cuda")
        padded[:, :n_genes] = torch.from_numpy(ranks.view(np.int32)).cuda()
        got = plan.from_rankings(padded[:, :n_genes]).cpu().numpy()
        got2 = plan.from_rankings(torch.from_numpy(ranks.view(np.int32)).cuda()).cpu().numpy()
    assert np.array_equal(got, got2)
    if unweighted:
        assert np.array_equal(got, want)
    else:
        assert_auc_close(got, want)


@pytest.mark.parametrize("unweighted", [False, True])
def test_uint16_column_indices_device_and_host_paths(unweighted, monkeypatch):
    """aucell_csr16_f32 (uint16 column indices) and the host pipeline that narrows int32 indices on the fly must give
    the bits of the int32 path; select-mode rows (> 4096 entries), empty rows and an out-of-range index included."""
    import torch

    n_cells, n_genes = 3000, 30000
    indptr, indices, data = synth.synth_csr_torch(n_cells, n_genes, 1800, seed=31, device="cuda", sigma=0.9)
    ip = indptr.cpu().numpy()
    assert (np.diff(ip) > 4096).any() and (np.diff(ip) < 1499).any()
    rs = np.random.RandomState(15)
    reg = synth.make_regulons(rs, n_genes, 120, 10, 400, weighted=not unweighted, repressors=True)
    cutoff = O.derive_rank_cutoff(0.05, n_genes)
    perm = np.random.RandomState(42).permutation(n_genes)
    with _plan_for(n_genes, perm, reg, cutoff) as plan:
        ref = plan.aucell_csr(indptr, indices, data)
        got16 = plan.aucell_csr(indptr, indices.to(torch.uint16), data)
        assert torch.equal(ref, got16)
        hi, hx, hd = ip, indices.cpu().numpy(), data.cpu().numpy()
        monkeypatch.setenv("AUCELL_B200_HOST_IDX16", "1")
        monkeypatch.setenv("AUCELL_B200_HOST_THREADS", "3")
        h16, nnz16 = plan.aucell_csr_host(hi, hx, hd, want_nnz=True, chunk_cells=257)
        monkeypatch.setenv("AUCELL_B200_HOST_IDX16", "0")
        h32, nnz32 = plan.aucell_csr_host(hi, hx, hd, want_nnz=True, chunk_cells=257)
        assert np.array_equal(h16, ref.cpu().numpy()) and np.array_equal(h32, h16) and np.array_equal(nnz16, nnz32)
        # a chunk large enough for the multi-threaded conversion, default chunking
        monkeypatch.setenv("AUCELL_B200_HOST_IDX16", "1")
        assert np.array_equal(plan.aucell_csr_host(hi, hx, hd), h16)
        # an index that does not fit 16 bits is refused, not truncated
        bad = hx.copy()
        bad[len(bad) // 2] = 70000
        with pytest.raises(Exception, match="out of range"):
            plan.aucell_csr_host(hi, bad, hd)
    X = synth.csr_to_dense(ip, indices.cpu().numpy(), data.cpu().numpy(), n_genes, 0, 64)
    assert_auc_close(h16[:64], oracle_auc(X, perm, reg, cutoff))


@pytest.mark.parametrize("thr,unweighted", [(0.05, False), (0.05, True), (0.97, True)])
def test_fused_more_than_65535_genes(thr, unweighted):
    """n_genes > 65535: 32-bit positions; thr 0.97 pushes the cutoff above 65535, where unit-weight sums no longer fit
    32 bits (64-bit accumulators with unit weights) and zeros / negatives fall below the cutoff."""
    import scipy.sparse as sp

    rs = np.random.RandomState(13)
    n_cells, n_genes = 12, 70001
    X = synth.dense_counts(rs, n_cells, n_genes, density=0.04)
    X[1] = 0.0
    X[2, rs.choice(n_genes, 3000, replace=False)] = -1.0
    X[3] = rs.randn(n_genes).astype(np.float32)
    reg = synth.make_regulons(rs, n_genes, 25, 10, 2000, weighted=True)
    cutoff = O.derive_rank_cutoff(thr, n_genes)
    perm = np.random.RandomState(9).permutation(n_genes)
    want = oracle_auc(X, perm, reg, cutoff, unweighted)
    csr = sp.csr_matrix(X)
    with _plan_for(n_genes, perm, reg, cutoff, unweighted) as plan:
        got_d = plan.aucell_dense(cuda(X)).cpu().numpy()
        got_s = plan.aucell_csr(cuda(csr.indptr.astype(np.int64)), cuda(csr.indices.astype(np.int32)),
                                cuda(csr.data)).cpu().numpy()
    assert np.array_equal(got_d, got_s)
    if unweighted:
        assert np.array_equal(got_d, want)
    else:
        assert_auc_close(got_d, want)


def test_wide_weight_range_uses_second_accumulator():
    """Weights spanning 14 orders of magnitude inside one regulon: the fixed-point scatter adds a second accumulator for
    the residual, so cells that only hit the tiny-weight genes still meet the 1e-6 bar."""
    rs = np.random.RandomState(3)
    n_cells, n_genes = 32, 2000
    X = synth.dense_counts(rs, n_cells, n_genes, density=0.1)
    names, indptr, genes, weights = synth.make_regulons(rs, n_genes, 12, 30, 200, weighted=True)
    weights = weights.copy()
    for k in range(len(names)):
        seg = slice(indptr[k], indptr[k + 1])
        w = weights[seg]
        w[::3] *= 1e-14                      # tiny weights next to O(1) weights
        w[1::7] *= 1e9
        weights[seg] = w
    reg = (names, indptr, genes, weights)
    cutoff = O.derive_rank_cutoff(0.05, n_genes)
    perm = np.random.RandomState(2).permutation(n_genes)
    # make some cells express ONLY tiny-weight genes of regulon 0
    tiny = genes[indptr[0]:indptr[1]][::3]
    others = np.setdiff1d(np.arange(n_genes), genes[indptr[0]:indptr[1]])[:300]   # fill the ranks below the cutoff
    X[0] = 0.0; X[0, tiny] = 3.0; X[0, others] = 1.0
    X[1] = 0.0; X[1, tiny[:5]] = 1.0; X[1, others] = 2.0
    want = oracle_auc(X, perm, reg, cutoff)
    with _plan_for(n_genes, perm, reg, cutoff) as plan:
        got = plan.aucell_dense(cuda(X)).cpu().numpy()
    assert want[0, 0] > 0 and want[0, 0] < 1e-10
    assert_auc_close(got, want)


def test_recovery_aucs_matches_oracle_on_a_feature_ranking_db():
    """ctxcore.recovery.aucs on a features x genes ranking database (the cisTarget use of the same arithmetic,
    reference transform.py:195): ranks refer to the whole genome, the frame holds only the module's genes."""
    from pyscenic_b200.recovery import aucs

    rs = np.random.RandomState(17)
    total_genes, n_feat, n_sel = 22284, 300, 120
    full = np.stack([rs.permutation(total_genes) for _ in range(n_feat)]).astype(np.int32)   # motif rankings
    sel = rs.choice(total_genes, n_sel, replace=False)
    rnk = pd.DataFrame(full[:, sel], index=["motif%d" % i for i in range(n_feat)], columns=["g%d" % g for g in sel])
    for w in (np.ones(n_sel), rs.lognormal(0, 1, n_sel)):
        for thr in (0.005, 0.05):
            got = aucs(rnk, total_genes, w, thr)
            want = O.aucs(rnk, total_genes, w, thr)
            if np.all(w == 1.0):
                assert np.array_equal(got, want)
            else:
                assert_auc_close(got.reshape(-1, 1), want.reshape(-1, 1))


@pytest.mark.parametrize("n_regulons,thr", [(2000, 0.02), (1200, 0.2)])
def test_regulon_count_sweep_vs_oracle(n_regulons, thr, fused_variant):
    """BASELINE.json configs[4]: regulon-count sweep (more regulons than threads per CTA) and a deep cutoff."""
    n_cells, n_genes = 40, 20000
    indptr, indices, data = synth.synth_csr_torch(n_cells, n_genes, 1400, seed=33, device="cuda", sigma=0.5)
    rs = np.random.RandomState(4)
    reg = synth.make_regulons(rs, n_genes, n_regulons, 10, 300, weighted=True)
    cutoff = O.derive_rank_cutoff(thr, n_genes)
    perm = np.random.RandomState(6).permutation(n_genes)
    X = synth.csr_to_dense(indptr.cpu().numpy(), indices.cpu().numpy(), data.cpu().numpy(), n_genes)
    want = oracle_auc(X, perm, reg, cutoff)
    with _plan_for(n_genes, perm, reg, cutoff) as plan:
        got = plan.aucell_csr(indptr, indices, data).cpu().numpy()
    assert_auc_close(got, want)


def test_c_abi_error_paths():
    """Invalid arguments are refused with an error code + message (no silent zeros)."""
    import ctypes

    import torch

    from pyscenic_b200 import _native
    from pyscenic_b200.plan import AUCellPlan, RegulonTables

    lib = _native.load()
    n_genes = 100
    perm = np.random.RandomState(0).permutation(n_genes)
    bad_perm = perm.copy(); bad_perm[3] = bad_perm[4]                       # not a permutation
    with pytest.raises(_native.AUCellNativeError, match="not a permutation"):
        AUCellPlan(n_genes, bad_perm, None)
    with pytest.raises(_native.AUCellNativeError):
        AUCellPlan(n_genes, perm, None, device=999)                          # no such device
    t = RegulonTables(names=["r"], indptr=np.array([0, 2], dtype=np.int64), cols=np.array([1, 500], dtype=np.int32),
                      weights=None, max_auc=np.array([10.0]), valid=np.ones(1, dtype=np.uint8), rank_cutoff=5, n_genes=n_genes)
    with pytest.raises(_native.AUCellNativeError, match="out of range"):
        AUCellPlan(n_genes, perm, t)                                         # gene column beyond the matrix
    t.cols = np.array([1, 2], dtype=np.int32); t.max_auc = np.array([0.0])
    with pytest.raises(_native.AUCellNativeError, match="max_auc"):
        AUCellPlan(n_genes, perm, t)
    t.max_auc = np.array([10.0]); t.weights = np.array([1.0, np.inf])
    with pytest.raises(_native.AUCellNativeError, match="finite"):
        AUCellPlan(n_genes, perm, t)
    # a ranking-only plan refuses AUC calls; NULL outputs are refused
    with AUCellPlan(n_genes, perm, None) as plan:
        x = torch.zeros((2, n_genes), dtype=torch.float32, device="cuda")
        with pytest.raises(_native.AUCellNativeError, match="no regulons"):
            plan.aucell_dense(x)
        rc = lib.aucell_rank_dense_f32(plan._handle, x.data_ptr(), 2, n_genes, None, None, None)
        assert rc == -1 and b"out_rank" in lib.aucell_b200_last_error()
        rc = lib.aucell_rank_dense_f32(plan._handle, x.data_ptr(), 2, n_genes - 1, x.data_ptr(), None, None)
        assert rc == -1                                                       # row_stride < n_genes
        assert lib.aucell_rank_dense_f32(plan._handle, None, 0, n_genes, None, None, None) == 0    # empty input is fine
    assert lib.aucell_b200_max_genes() >= 500_000
    assert lib.aucell_b200_launch_count() > 0


def test_randomised_differential_fuzz_short():
    """A short run of scripts/fuzz_gpu.py (random shapes, densities, value distributions, thresholds, regulon sets; every
    entry point against the oracle).  The round-1 long run covered 9,253 cases (profiles/r1_fuzz_gpu.txt)."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "scripts", "fuzz_gpu.py"), "12", "7"], cwd=root,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0 and "fuzz ok" in res.stdout, res.stdout[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize("unweighted", [False, True])
def test_dense_host_call_squeezes_zeros_out_on_the_host(unweighted, monkeypatch):
    """aucell_dense_f32_host on a sparse dense matrix: the host threads drop the zeros (AVX2 compress) and the call
    continues as a CSR call.  Bit-identical to the dense device path and to the CSR entry point -- with -0.0, NaN, +inf,
    negatives, an all-zero row and a row stride larger than the gene count in the matrix -- and switched off
    (AUCELL_B200_HOST_COMPACT=0, or a dense matrix) the result is the same."""
    import scipy.sparse as sp

    n_cells, n_genes, stride = 800, 6011, 6016
    rs = np.random.RandomState(5)
    X = synth.dense_counts(rs, n_cells, n_genes, density=0.08)
    X[3] = 0.0
    X[5, rs.choice(n_genes, 40, replace=False)] = -1.5
    X[6, 17] = np.nan
    X[7, 23] = np.inf
    X[8, rs.choice(n_genes, 100, replace=False)] = -0.0
    X[9, -9:] = 3.0                                       # entries in the scalar remainder of the row
    buf = np.full((n_cells, stride), 7.0, dtype=np.float32)          # (the padding must never be read as genes)
    buf[:, :n_genes] = X
    view = buf[:, :n_genes]
    reg = synth.make_regulons(rs, n_genes, 80, 10, 300, weighted=not unweighted)
    cutoff = O.derive_rank_cutoff(0.05, n_genes)
    perm = np.random.RandomState(3).permutation(n_genes)
    m = sp.csr_matrix(X)
    from pyscenic_b200 import _native

    lib = _native.load()
    with _plan_for(n_genes, perm, reg, cutoff, unweighted=unweighted) as plan:
        want = plan.aucell_dense(cuda(X)).cpu().numpy()
        want_csr = plan.aucell_csr(cuda(m.indptr.astype(np.int64)), cuda(m.indices.astype(np.int32)), cuda(m.data)).cpu().numpy()
        assert np.array_equal(want, want_csr)

        def host_call(arr, row_stride):
            out = np.full((n_cells, len(reg[0])), -1.0)
            nnz = np.full(n_cells, -1, dtype=np.int32)
            _native.check(lib.aucell_dense_f32_host(plan._handle, arr.ctypes.data, n_cells, row_stride, out.ctypes.data,
                                                    nnz.ctypes.data, 0))
            return out, nnz

        got, nnz = host_call(buf, stride)                 # squeezed on the host (n_cells * n_genes >= 2^22, 8 % dense)
        assert np.array_equal(got, want)
        assert np.array_equal(nnz, np.count_nonzero(view, axis=1))
        monkeypatch.setenv("AUCELL_B200_HOST_COMPACT", "0")
        got0, nnz0 = host_call(buf, stride)               # uploaded as it is
        assert np.array_equal(got0, want) and np.array_equal(nnz0, nnz)
        monkeypatch.delenv("AUCELL_B200_HOST_COMPACT")
        # the layouts and float types a DataFrame block comes in (aucell_dense_strided_host)
        for dt in (np.float32, np.float64):
            for order in ("C", "F"):
                arr = np.asarray(X, dtype=dt, order=order)
                out = np.full_like(want, -1.0)
                assert plan.aucell_dense_strided_host(arr, out), (dt, order)
                assert np.array_equal(out, want), (dt, order)
        gm = np.asfortranarray(buf)[:, :n_genes]          # gene-major with a gene stride larger than the cell count
        gm_rows = gm[100:]
        out = np.full((n_cells - 100, want.shape[1]), -1.0)
        assert plan.aucell_dense_strided_host(gm_rows, out) and np.array_equal(out, want[100:])
        inexact = X.astype(np.float64)
        inexact[11, 5] = 0.1                              # not a float32 value: the cast would change it
        assert not plan.aucell_dense_strided_host(inexact, np.empty_like(want))
        assert not plan.aucell_dense_strided_host(np.asfortranarray(inexact), np.empty_like(want))
    Xd = np.where(rs.random_sample((n_cells, n_genes)) < 0.6, X.max() + rs.random_sample((n_cells, n_genes)), 0).astype(np.float32)
    with _plan_for(n_genes, perm, reg, cutoff, unweighted=unweighted) as plan:       # 60 % dense: stays dense
        want_d = plan.aucell_dense(cuda(Xd)).cpu().numpy()
        out = np.full((n_cells, len(reg[0])), -1.0)
        _native.check(lib.aucell_dense_f32_host(plan._handle, Xd.ctypes.data, n_cells, n_genes, out.ctypes.data, None, 0))
        assert np.array_equal(out, want_d)


@pytest.mark.gpu
def test_malformed_device_csr_is_memory_safe():
    """Device-side CSR input is not validated by a separate pass (the Python wrapper checks scipy input on the host):
    a column index outside [0, n_genes) must not be used as an address -- the tie-class kernels hand such a cell over
    (the bit count does not match) and the general kernel drops the entry, i.e. ranks the cell as if it were zero."""
    import scipy.sparse as sp
    import torch

    n_cells, n_genes = 600, 5003
    rs = np.random.RandomState(17)
    X = synth.dense_counts(rs, n_cells, n_genes, 0.08)
    m = sp.csr_matrix(X)
    reg = synth.make_regulons(rs, n_genes, 60, 10, 300, weighted=True)
    cutoff = O.derive_rank_cutoff(0.05, n_genes)
    perm = np.random.RandomState(3).permutation(n_genes)
    bad_idx = m.indices.astype(np.int32).copy()
    fixed = X.copy()
    for r, col in ((5, n_genes), (77, n_genes + 12345), (300, 2_000_000_000), (451, -7)):
        e = m.indptr[r] + 3
        fixed[r, m.indices[e]] = 0.0                      # the entry the kernels must ignore
        bad_idx[e] = col
    good = sp.csr_matrix(fixed)
    with _plan_for(n_genes, perm, reg, cutoff) as plan:
        got = plan.aucell_csr(cuda(m.indptr.astype(np.int64)), cuda(bad_idx), cuda(m.data)).cpu().numpy()
        want = plan.aucell_csr(cuda(good.indptr.astype(np.int64)), cuda(good.indices.astype(np.int32)), cuda(good.data)).cpu().numpy()
        fs = plan.fast_stats()
        plan.set_fast_path(False)
        got_general = plan.aucell_csr(cuda(m.indptr.astype(np.int64)), cuda(bad_idx), cuda(m.data)).cpu().numpy()
    assert fs["handed_duplicate_column"] >= 4
    assert np.array_equal(got, want) and np.array_equal(got_general, want)
