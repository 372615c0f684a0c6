"""Randomised differential test: the CUDA path (C-ABI, every input kind / dtype) against the CPU oracle on random shapes,
densities, value distributions, thresholds and regulon sets.  python scripts/fuzz_gpu.py [seconds] [seed]"""
import sys, time
sys.path.insert(0, ".")
import numpy as np, scipy.sparse as sp, torch
import synth
from oracle import aucell_oracle as O
from pyscenic_b200.plan import AUCellPlan, derive_rank_cutoff

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cuda = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
t_end = time.time() + budget
it = 0; worst = 0.0
while time.time() < t_end:
    rs = np.random.RandomState(seed0 * 100003 + it); it += 1
    n_genes = int(rs.choice([2, 3, 17, 64, 257, 1000, 2049, 4097, 6000, 9001]))
    n_cells = int(rs.randint(1, 24))
    kind = rs.randint(0, 6)
    dens = float(rs.choice([0.0, 0.01, 0.05, 0.2, 0.6, 1.0]))
    if kind == 0:   X = np.where(rs.rand(n_cells, n_genes) < dens, rs.poisson(1.5, (n_cells, n_genes)) + 1, 0).astype(np.float64)
    elif kind == 1: X = np.where(rs.rand(n_cells, n_genes) < dens, np.log1p(rs.geometric(0.4, (n_cells, n_genes))), 0.0)
    elif kind == 2: X = np.where(rs.rand(n_cells, n_genes) < dens, rs.randn(n_cells, n_genes), 0.0)
    elif kind == 3: X = np.where(rs.rand(n_cells, n_genes) < dens, np.round(rs.randn(n_cells, n_genes) * 2) / 2, 0.0)      # few values, both signs
    elif kind == 4: X = np.where(rs.rand(n_cells, n_genes) < dens, 1.0 + rs.randint(0, 50, (n_cells, n_genes)) * 1e-7, 0.0)  # crowded bins
    else:           X = np.where(rs.rand(n_cells, n_genes) < dens, rs.lognormal(0, 2, (n_cells, n_genes)), 0.0)
    if rs.rand() < 0.3 and n_genes > 3:
        X[rs.randint(n_cells), rs.randint(n_genes)] = np.nan
        X[rs.randint(n_cells), rs.randint(n_genes)] = np.inf
        X[rs.randint(n_cells), rs.randint(n_genes)] = -0.0
    use64 = rs.rand() < 0.3
    X = X.astype(np.float64 if use64 else np.float32)
    perm = rs.permutation(n_genes)
    # rankings
    want_r = O.create_rankings_c(X, perm)
    with AUCellPlan(n_genes, perm, None) as plan:
        got_r = plan.rank_dense(cuda(X)).cpu().numpy().view(np.uint32)
        csr = sp.csr_matrix(X)
        got_rs = plan.rank_csr(cuda(csr.indptr.astype(np.int64)), cuda(csr.indices.astype(np.int32)), cuda(csr.data)).cpu().numpy().view(np.uint32)
    assert np.array_equal(got_r, want_r), ("rank dense", it, n_genes, kind, dens, use64)
    assert np.array_equal(got_rs, want_r), ("rank csr", it, n_genes, kind, dens, use64)
    # AUCs
    if n_genes < 17: continue
    thr = float(rs.choice([0.01, 0.05, 0.2, 0.5, 0.9]))
    try: cutoff = derive_rank_cutoff(thr, n_genes)
    except AssertionError: continue
    R = int(rs.choice([1, 7, 60, 600]))
    reg = synth.make_regulons(rs, n_genes, R, 2, min(n_genes, 300), weighted=True)
    unweighted = rs.rand() < 0.4
    names, indptr, genes, weights = reg
    inv = np.empty_like(perm); inv[perm] = np.arange(n_genes)
    cols, ws = [], []
    for k in range(R):
        p = inv[genes[indptr[k]:indptr[k + 1]]]; o = np.argsort(p, kind="stable")
        cols.append(p[o]); ws.append(np.ones(len(o)) if unweighted else weights[indptr[k]:indptr[k + 1]][o])
    want = O.auc_matrix_from_ranks(want_r, cols, ws, [True] * R, cutoff)
    tables = synth.regulon_tables_from_arrays(*reg, n_genes=n_genes, rank_cutoff=cutoff, unweighted=unweighted)
    with AUCellPlan(n_genes, perm, tables) as plan:
        got_d = plan.aucell_dense(cuda(X)).cpu().numpy()
        got_s = plan.aucell_csr(cuda(csr.indptr.astype(np.int64)), cuda(csr.indices.astype(np.int32)), cuda(csr.data)).cpu().numpy()
        if not use64:     # uint16 column indices (device entry point) and the host pipelines (int32 kept / narrowed)
            got_16 = plan.aucell_csr(cuda(csr.indptr.astype(np.int64)), cuda(csr.indices.astype(np.uint16)), cuda(csr.data)).cpu().numpy()
            got_h = plan.aucell_csr_host(csr.indptr.astype(np.int64), csr.indices.astype(np.int32), csr.data, chunk_cells=int(rs.randint(1, 9)))
            got_hd = plan.aucell_dense_host(X, chunk_cells=int(rs.randint(1, 9)))
            assert np.array_equal(got_16, got_s), ("csr16 vs csr", it, n_genes, kind, dens, thr, R)
            assert np.array_equal(got_h, got_s), ("csr host vs csr", it, n_genes, kind, dens, thr, R)
            assert np.array_equal(got_hd, got_s), ("dense host vs csr", it, n_genes, kind, dens, thr, R)
    with AUCellPlan(n_genes, None, synth.regulon_tables_from_arrays(names, indptr, inv[genes].astype(np.int32), weights, n_genes=n_genes, rank_cutoff=cutoff, unweighted=unweighted)) as plan4r:
        got_4r = plan4r.from_rankings(cuda(want_r.view(np.int32))).cpu().numpy()
    assert np.array_equal(got_d, got_s), ("dense vs csr", it, n_genes, kind, dens, thr, R)
    assert np.array_equal(got_d, got_4r), ("fused vs from_rankings", it, n_genes, kind, dens, thr, R)
    if unweighted:
        assert np.array_equal(got_d, want), ("unweighted", it, n_genes, kind, dens, thr, R)
    else:
        assert np.array_equal(got_d == 0, want == 0), ("zero pattern", it)
        err = np.abs(got_d - want) / np.maximum(np.abs(want), 1e-300); err[want == 0] = 0
        worst = max(worst, float(err.max()))
        assert err.max() <= 1e-6, ("weighted", it, n_genes, kind, dens, thr, R, err.max())
print("fuzz ok: %d random cases, worst weighted relative error %.3g" % (it, worst))
