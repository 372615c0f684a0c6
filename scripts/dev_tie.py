"""Development check of the tie-class kernel: fast path on/off must agree bit for bit; timing of both."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import synth
from pyscenic_b200.plan import AUCellPlan, derive_rank_cutoff

n_cells = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
n_genes, mean_nnz, R = 27_998, 1960, 500
for weighted in (True, False):
    rs = np.random.RandomState(1234)
    reg = synth.make_regulons(rs, n_genes, R, 10, 400, weighted=weighted, repressors=True)
    cutoff = derive_rank_cutoff(0.05, n_genes)
    tables = synth.regulon_tables_from_arrays(*reg, n_genes=n_genes, rank_cutoff=cutoff, unweighted=not weighted)
    perm = np.random.RandomState(42).permutation(n_genes)
    plan = AUCellPlan(n_genes, perm, tables, device=0)
    dev = torch.device("cuda", 0)
    indptr, indices, data = synth.synth_csr_torch(n_cells, n_genes, mean_nnz, seed=1000, device=dev)
    res = {}
    for fast in (2, 1, False):                      # tie-class kernel version (AUCELL_B200_TIE_KERNEL), or the general kernel
        if fast:
            os.environ["AUCELL_B200_TIE_KERNEL"] = str(fast)
        avail = plan.set_fast_path(bool(fast))
        out = torch.full((n_cells, R), -1.0, dtype=torch.float64, device=dev)
        nnz = torch.full((n_cells,), -1, dtype=torch.int32, device=dev)
        for _ in range(2):
            plan.aucell_csr(indptr, indices, data, out=out, out_nnz=nnz)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            plan.aucell_csr(indptr, indices, data, out=out, out_nnz=nnz)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        res[fast] = (out.clone(), nnz.clone())
        print(json.dumps({"weighted": weighted, "fast": fast, "available": avail, "ms": ms,
                          "Mcells_per_s": n_cells / ms / 1e3, "fast_stats": plan.fast_stats(),
                          "path_stats": plan.path_stats()}), flush=True)
    for v in (2, 1):
        same = torch.equal(res[v][0], res[False][0]); same_nnz = torch.equal(res[v][1], res[False][1])
        diff = (res[v][0] - res[False][0]).abs().max().item()
        print("weighted", weighted, "kernel", v, "identical:", same, "nnz identical:", same_nnz, "max abs diff", diff, flush=True)
        if not same:
            bad = (res[v][0] != res[False][0]).any(dim=1).nonzero().flatten()
            print("bad cells:", bad.numel(), bad[:10].tolist())
            c = int(bad[0]); cols = (res[v][0][c] != res[False][0][c]).nonzero().flatten()[:5].tolist()
            print("cell", c, "nnz", int(indptr[c+1]-indptr[c]), "cols", cols, res[v][0][c, cols].tolist(), res[False][0][c, cols].tolist())
    plan.close()
