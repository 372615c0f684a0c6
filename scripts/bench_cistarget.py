"""cisTarget AUC step on the AUCell scatter kernel: features x genes ranking database, all modules
in one launch.  Prints one JSON line: GPU kernel rate (features x modules per second), HBM GB/s, and the CPU oracle
(ctxcore.recovery.aucs restated, one module at a time as transform.py:175-195 does) on a bounded sample."""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np, pandas as pd, torch
import synth
from oracle import aucell_oracle as O
from pyscenic_b200.plan import AUCellPlan, derive_rank_cutoff

n_feat, n_genes, n_mod = 24453, 22284, 2000            # motif collection v9 x hg19 genes, modules of one SCENIC run
g = torch.Generator(device="cuda"); g.manual_seed(5)
ranks = torch.empty((n_feat, n_genes), dtype=torch.int32, device="cuda")
for s in range(0, n_feat, 2048):
    e = min(n_feat, s + 2048)
    ranks[s:e] = torch.argsort(torch.rand((e - s, n_genes), generator=g, device="cuda"), dim=1).to(torch.int32)
rs = np.random.RandomState(2)
reg = synth.make_regulons(rs, n_genes, n_mod, 20, 1500, weighted=True)
cutoff = derive_rank_cutoff(0.05, n_genes)
tables = synth.regulon_tables_from_arrays(*reg, n_genes=n_genes, rank_cutoff=cutoff)
out = {"workload": "%d features x %d genes int32 ranks, %d weighted modules (20-1500 genes), thr 0.05" % (n_feat, n_genes, n_mod)}
with AUCellPlan(n_genes, None, tables) as plan:
    auc = torch.empty((n_feat, n_mod), dtype=torch.float64, device="cuda")
    fn = lambda: plan.from_rankings(ranks, out=auc)
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [fn() for _ in range(5)]; e1.record(); torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) / 1e3 / 5
    out["gpu_seconds_all_modules"] = sec
    out["gpu_feature_module_aucs_per_s"] = n_feat * n_mod / sec
    out["gpu_GBps"] = n_feat * (4 * n_genes + 8 * n_mod) / sec / 1e9
    nes = (auc - auc.mean(dim=0)) / auc.std(dim=0, unbiased=False)
    out["enriched_nes_ge_3"] = int((nes >= 3.0).sum().item())
# CPU: the reference's per-module loop on a sample of modules (each: column gather + numba-equivalent C loop)
names, indptr, genes, weights = reg
rk = ranks.cpu().numpy()
t0 = time.perf_counter(); n_s = 20
mx = 0.0
for m in range(n_s):
    c = np.sort(genes[indptr[m]:indptr[m + 1]]); w = weights[indptr[m]:indptr[m + 1]][np.argsort(genes[indptr[m]:indptr[m + 1]], kind="stable")]
    exp = O.aucs(pd.DataFrame(rk[:, c]), n_genes, w, 0.05)
    mx = max(mx, float(np.max(np.abs(auc[:, m].cpu().numpy() - exp) / np.maximum(exp, 1e-300))))
cpu = (time.perf_counter() - t0) / n_s
out["cpu_seconds_per_module"] = cpu
out["cpu_feature_module_aucs_per_s"] = n_feat / cpu
out["max_rel_err_vs_oracle"] = mx
out["speedup_vs_cpu_1core"] = out["gpu_feature_module_aucs_per_s"] / out["cpu_feature_module_aucs_per_s"]
print(json.dumps(out))
