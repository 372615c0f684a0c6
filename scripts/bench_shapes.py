"""Throughput of the fused kernel away from the headline shape: long rows (select mode), float64 values, very sparse
cells, > 65535 genes.  One JSON line per case with the path statistics."""
import sys, json
sys.path.insert(0, ".")
import numpy as np, torch
import synth
from pyscenic_b200.plan import AUCellPlan, derive_rank_cutoff

import os
ONLY = os.environ.get("ONLY", "")


def run(name, n_cells, n_genes, mean_nnz, R=500, f64=False, sigma=0.35, thr=0.05):
    if ONLY and ONLY not in name:
        return
    rs = np.random.RandomState(1)
    reg = synth.make_regulons(rs, n_genes, R, 10, 400, weighted=True)
    cutoff = derive_rank_cutoff(thr, n_genes)
    tables = synth.regulon_tables_from_arrays(*reg, n_genes=n_genes, rank_cutoff=cutoff)
    perm = np.random.RandomState(42).permutation(n_genes)
    indptr, indices, data = synth.synth_csr_torch(n_cells, n_genes, mean_nnz, seed=2, device="cuda", sigma=sigma)
    if f64: data = data.double()
    with AUCellPlan(n_genes, perm, tables) as plan:
        out = torch.empty((n_cells, R), dtype=torch.float64, device="cuda")
        plan.path_stats()
        plan.aucell_csr(indptr, indices, data, out=out); torch.cuda.synchronize()
        st = plan.path_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); [plan.aucell_csr(indptr, indices, data, out=out) for _ in range(3)]; e1.record(); torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) / 1e3 / 3
    nnz = int(indptr[-1].item())
    print(json.dumps({"case": name, "cells_per_s": n_cells / sec, "GBps": (nnz * (12 if f64 else 8) + 8 * R * n_cells) / sec / 1e9,
                      "nnz_per_cell": nnz / n_cells, "path_stats": st}))

run("cfg4 shape, float32 (reference point)", 100000, 27998, 1960)
run("cfg4 shape, float64 values", 100000, 27998, 1960, f64=True)
run("long rows: 8,400 stored entries per cell (30 % of 27,998)", 40000, 27998, 8400)
run("very sparse cells: 60 stored entries per cell", 400000, 27998, 60)
run("70,001 genes (32-bit positions), 3,500 entries per cell", 60000, 70001, 3500)
run("deep cells: 6,000 entries per cell of 60,000 genes (selection pruned by position)", 40000, 60000, 6000)
