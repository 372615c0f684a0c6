"""Short run of the tie-class kernel for ncu: weighted cfg4 shape, a few launches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import synth
from pyscenic_b200.plan import AUCellPlan, derive_rank_cutoff

n_cells = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000
weighted = (sys.argv[2] != "unit") if len(sys.argv) > 2 else True
n_genes, mean_nnz, R = 27_998, 1960, 500
rs = np.random.RandomState(1234)
reg = synth.make_regulons(rs, n_genes, R, 10, 400, weighted=weighted, repressors=True)
cutoff = derive_rank_cutoff(0.05, n_genes)
tables = synth.regulon_tables_from_arrays(*reg, n_genes=n_genes, rank_cutoff=cutoff, unweighted=not weighted)
plan = AUCellPlan(n_genes, np.random.RandomState(42).permutation(n_genes), tables, device=0)
dev = torch.device("cuda", 0)
indptr, indices, data = synth.synth_csr_torch(n_cells, n_genes, mean_nnz, seed=1000, device=dev)
out = torch.empty((n_cells, R), dtype=torch.float64, device=dev)
for _ in range(3):
    plan.aucell_csr(indptr, indices, data, out=out)
torch.cuda.synchronize()
print("ok", float(out.sum()))
