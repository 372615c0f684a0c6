"""Where the wall time of aucell(scipy_csr) / aucell(DataFrame) goes: wraps the host-side phases with timers."""
import sys, time, json, functools
sys.path.insert(0, ".")
import numpy as np, pandas as pd, scipy.sparse as sp, torch
import synth
import pyscenic_b200.aucell as A
import pyscenic_b200.plan as P

T = {}
def timed(mod, name):
    fn = getattr(mod, name)
    @functools.wraps(fn)
    def w(*a, **k):
        t0 = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            T[name] = T.get(name, 0.0) + time.perf_counter() - t0
    setattr(mod, name, w)

for m, n in ((A, "permutation_from_seed"), (A, "prepare_regulons"), (A, "_canonical_csr"), (A, "_assemble"), (A, "_csr_auc_shard"),
             (A, "_dense_auc_shard")):
    timed(m, n)
_init = P.AUCellPlan.__init__
def init(self, *a, **k):
    t0 = time.perf_counter(); _init(self, *a, **k); T["plan_create"] = T.get("plan_create", 0.0) + time.perf_counter() - t0
P.AUCellPlan.__init__ = init
_close = P.AUCellPlan.close
def close(self):
    t0 = time.perf_counter(); _close(self); T["plan_close"] = T.get("plan_close", 0.0) + time.perf_counter() - t0
P.AUCellPlan.close = close

n_cells, n_genes, R = int(sys.argv[1]) if len(sys.argv) > 1 else 300000, 27998, 500
indptr, indices, data = synth.synth_csr_torch(n_cells, n_genes, 1960, seed=3, device="cuda")
csr = sp.csr_matrix((data.cpu().numpy(), indices.cpu().numpy(), indptr.cpu().numpy()), shape=(n_cells, n_genes))
del indptr, indices, data
rs = np.random.RandomState(1)
reg = synth.make_regulons(rs, n_genes, R, 10, 400, weighted=True)
sigs = synth.regulons_to_signatures(*reg)
genes = ["g%d" % j for j in range(n_genes)]
A.aucell(csr[:2048], sigs, seed=42, genes=genes)
for rep in range(2):
    T.clear()
    t0 = time.perf_counter(); out = A.aucell(csr, sigs, seed=42, genes=genes); dt = time.perf_counter() - t0
    print(json.dumps({"call": "aucell(csr)", "cells": n_cells, "seconds": dt, "cells_per_s": n_cells / dt, "phases": dict(T)}), flush=True)
n_df = 20000
for dt_ in (np.float32, np.float64):
    for order in ("C", "F"):
        X = np.asarray(csr[:n_df].toarray(), dtype=dt_, order=order)
        df = pd.DataFrame(X, index=["c%d" % i for i in range(n_df)], columns=genes, copy=False)
        A.aucell(df.iloc[:256], sigs, seed=42)
        T.clear()
        t0 = time.perf_counter(); out = A.aucell(df, sigs, seed=42); dt = time.perf_counter() - t0
        print(json.dumps({"call": "aucell(DataFrame %s %s-order)" % (np.dtype(dt_).name, order), "cells": n_df, "seconds": dt,
                          "cells_per_s": n_df / dt, "phases": dict(T)}), flush=True)
