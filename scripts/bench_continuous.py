"""Fused kernel on inputs that are NOT count-like: dense continuous matrices (z-scored / imputed expression), every
value distinct, half of them negative; and a heavy-tailed positive matrix.  Reports cells/s and which path the cells
took (select mode streams the row three times; crowded value bins go to the general sort)."""
import sys, json
sys.path.insert(0, ".")
import numpy as np, torch
import synth
from pyscenic_b200.plan import AUCellPlan, derive_rank_cutoff

n_cells, n_genes, R = 8000, 20000, 300
rs = np.random.RandomState(3)
reg = synth.make_regulons(rs, n_genes, R, 10, 400, weighted=True)
cutoff = derive_rank_cutoff(0.05, n_genes)
tables = synth.regulon_tables_from_arrays(*reg, n_genes=n_genes, rank_cutoff=cutoff)
perm = np.random.RandomState(42).permutation(n_genes)
g = torch.Generator(device="cuda"); g.manual_seed(0)
cases = {
    "z-scored (randn, all distinct, half negative)": torch.randn((n_cells, n_genes), generator=g, device="cuda"),
    "log-normal sigma 2 (heavy tail, all positive)": torch.exp(2.0 * torch.randn((n_cells, n_genes), generator=g, device="cuda")),
    "uniform (0,1) with 70% zeros": torch.rand((n_cells, n_genes), generator=g, device="cuda") * (torch.rand((n_cells, n_genes), generator=g, device="cuda") < 0.3),
}
with AUCellPlan(n_genes, perm, tables) as plan:
    for name, X in cases.items():
        X = X.contiguous().float()
        out = torch.empty((n_cells, R), dtype=torch.float64, device="cuda")
        plan.path_stats()
        plan.aucell_dense(X, out=out); torch.cuda.synchronize()
        st = plan.path_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); [plan.aucell_dense(X, out=out) for _ in range(3)]; e1.record(); torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) / 1e3 / 3
        print(json.dumps({"case": name, "cells_per_s": n_cells / sec, "GBps": n_cells * n_genes * 4 / sec / 1e9, "path_stats": st}))
