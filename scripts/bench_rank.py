"""Throughput of the create_rankings and aucell4r kernels (secondary paths), cells/s on one GPU."""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np, torch
import synth
from pyscenic_b200.plan import AUCellPlan, derive_rank_cutoff
n_cells, n_genes = 20000, 27998
perm = np.random.RandomState(42).permutation(n_genes)
indptr, indices, data = synth.synth_csr_torch(n_cells, n_genes, 1960, seed=3, device="cuda")
rs = np.random.RandomState(1)
reg = synth.make_regulons(rs, n_genes, 500, 10, 400, weighted=True)
cutoff = derive_rank_cutoff(0.05, n_genes)
tables = synth.regulon_tables_from_arrays(*reg, n_genes=n_genes, rank_cutoff=cutoff)
out = {}
with AUCellPlan(n_genes, perm, None) as plan:
    ranks = torch.empty((n_cells, n_genes), dtype=torch.int32, device="cuda")
    for name, fn in (("rank_csr", lambda: plan.rank_csr(indptr, indices, data, out=ranks)),):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); [fn() for _ in range(3)]; e1.record(); torch.cuda.synchronize()
        out[name + "_cells_per_s"] = 3 * n_cells / (e0.elapsed_time(e1) / 1e3)
    X = torch.zeros((2000, n_genes), dtype=torch.float32, device="cuda")
    ip = indptr[:2001].cpu().numpy(); ix = indices[: ip[-1]].cpu().numpy(); dv = data[: ip[-1]].cpu().numpy()
    X = torch.from_numpy(synth.csr_to_dense(ip, ix, dv, n_genes)).cuda()
    r2 = torch.empty((2000, n_genes), dtype=torch.int32, device="cuda")
    fn = lambda: plan.rank_dense(X, out=r2)
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [fn() for _ in range(3)]; e1.record(); torch.cuda.synchronize()
    out["rank_dense_cells_per_s"] = 3 * 2000 / (e0.elapsed_time(e1) / 1e3)
    assert torch.equal(r2, ranks[:2000])
with AUCellPlan(n_genes, None, tables) as plan4r:     # columns of the ranking matrix are already shuffled positions
    auc = torch.empty((n_cells, 500), dtype=torch.float64, device="cuda")
    fn = lambda: plan4r.from_rankings(ranks, out=auc)
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [fn() for _ in range(3)]; e1.record(); torch.cuda.synchronize()
    out["from_rankings_cells_per_s"] = 3 * n_cells / (e0.elapsed_time(e1) / 1e3)
    out["from_rankings_GBps"] = out["from_rankings_cells_per_s"] * (4 * n_genes + 8 * 500) / 1e9
out["rank_csr_write_GBps"] = out["rank_csr_cells_per_s"] * 4 * n_genes / 1e9
print(json.dumps(out))
