"""BASELINE.json configs[4]: auc_threshold sweep 0.01-0.2 and regulon-count sweep 50-2,000 on 200k cells x 20k genes CSR.
Prints one JSON line per point (cells/s, device-resident, CUDA events, 3 timed launches after 1 warm-up)."""
import json, sys
sys.path.insert(0, ".")
import numpy as np, torch
import synth
from pyscenic_b200.plan import AUCellPlan, derive_rank_cutoff

n_cells, n_genes = 200_000, 20_000
indptr, indices, data = synth.synth_csr_torch(n_cells, n_genes, 1400, seed=11, device="cuda")
perm = np.random.RandomState(42).permutation(n_genes)
out = torch.empty((n_cells, 2000), dtype=torch.float64, device="cuda")
points = [(thr, 500) for thr in (0.01, 0.02, 0.05, 0.1, 0.2)] + [(0.05, r) for r in (50, 100, 250, 1000, 2000)]
for thr, R in points:
    rs = np.random.RandomState(7)
    reg = synth.make_regulons(rs, n_genes, R, 10, 400, weighted=True)
    cutoff = derive_rank_cutoff(thr, n_genes)
    tables = synth.regulon_tables_from_arrays(*reg, n_genes=n_genes, rank_cutoff=cutoff)
    with AUCellPlan(n_genes, perm, tables) as plan:
        o = out[:, :R].contiguous() if R != 2000 else out
        plan.aucell_csr(indptr, indices, data, out=o); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): plan.aucell_csr(indptr, indices, data, out=o)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        st = plan.path_stats()
    nnz = float(indptr[-1]) / n_cells
    bytes_cell = 8 * nnz + 8 + 8 * R
    print(json.dumps({"auc_threshold": thr, "rank_cutoff": cutoff, "n_regulons": R, "cells_per_s": n_cells / (ms / 1e3),
                      "ms": ms, "GBps_algorithmic": n_cells * bytes_cell / (ms / 1e3) / 1e9,
                      "general_path_cells": st["general_too_many_entries"] + st["general_negatives"] + st["general_crowded_bin"],
                      "cells_total": st["cells"]}), flush=True)
