"""Regulon specificity scores: GPU kernel time on the cfg4-sized AUC matrix (1.3 M cells x 500
regulons, 30 cell types) and the reference's loop (scipy jensenshannon per (regulon, type)) on a bounded sample."""
import sys, time, json, ctypes
sys.path.insert(0, ".")
import numpy as np, torch
from scipy.spatial.distance import jensenshannon
from pyscenic_b200 import _native

n_cells, n_reg, n_types = 1_300_000, 500, 30
g = torch.Generator(device="cuda"); g.manual_seed(1)
auc = torch.rand((n_cells, n_reg), generator=g, device="cuda", dtype=torch.float64) * 0.1
codes = torch.randint(0, n_types, (n_cells,), generator=g, device="cuda", dtype=torch.int32)
counts = torch.bincount(codes.long(), minlength=n_types).cpu().numpy().astype(np.int64)
out = torch.empty((n_types, n_reg), dtype=torch.float64, device="cuda")
lib = _native.load()
fn = lambda: _native.check(lib.aucell_rss_f64(0, auc.data_ptr(), n_cells, n_reg, n_reg, codes.data_ptr(), n_types,
                                              counts.ctypes.data_as(ctypes.c_void_p), out.data_ptr(), None))
fn(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); [fn() for _ in range(5)]; e1.record(); torch.cuda.synchronize()
sec = e0.elapsed_time(e1) / 1e3 / 5
res = {"workload": "%d cells x %d regulons, %d cell types (fp64 AUC matrix, %.1f GB)" % (n_cells, n_reg, n_types, n_cells * n_reg * 8 / 1e9),
       "gpu_seconds": sec, "gpu_GBps": 2 * n_cells * n_reg * 8 / sec / 1e9}
a = auc[:, :4].cpu().numpy(); lab = codes.cpu().numpy()
t0 = time.perf_counter(); mx = 0.0
for r in range(4):
    for t in range(3):
        q = (lab == t).astype(int)
        v = 1.0 - jensenshannon(a[:, r] / a[:, r].sum(), q / q.sum())
        mx = max(mx, abs(v - float(out[t, r].item())))
cpu = (time.perf_counter() - t0) / 12
res.update({"cpu_seconds_per_pair": cpu, "cpu_seconds_all_pairs_extrapolated": cpu * n_reg * n_types,
            "max_abs_diff_vs_scipy": mx, "speedup_vs_cpu_1core": cpu * n_reg * n_types / sec})
print(json.dumps(res))
