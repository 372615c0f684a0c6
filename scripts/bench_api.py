"""Wall time of the drop-in Python call  aucell(DataFrame, signatures, seed=42)  (the call a pySCENIC user makes),
with its host-side phases, for float32 and float64 frames (pd.read_csv yields float64)."""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np, pandas as pd
import synth
from pyscenic_b200 import aucell as A
from pyscenic_b200.plan import permutation_from_seed, prepare_regulons

n_cells, n_genes, R = 20000, 20000, 300
rs = np.random.RandomState(0)
X = synth.dense_counts(rs, n_cells, n_genes, 0.10)
reg = synth.make_regulons(rs, n_genes, R, 10, 1000, weighted=True)
sigs = synth.regulons_to_signatures(*reg)
for dt in (np.float32, np.float64):
    df = pd.DataFrame(X.astype(dt), index=["c%d" % i for i in range(n_cells)], columns=["g%d" % j for j in range(n_genes)])
    A.aucell(df.iloc[:64], sigs, seed=42)                      # warm-up (library load, context)
    t0 = time.perf_counter(); out = A.aucell(df, sigs, seed=42); t1 = time.perf_counter()
    ta = time.perf_counter(); perm = permutation_from_seed(n_genes, 42); tables = prepare_regulons(df.columns, perm, sigs, 0.05); tb = time.perf_counter()
    v = A._dense_values(df.to_numpy()); tc = time.perf_counter()
    print(json.dumps({"dtype": np.dtype(dt).name, "cells": n_cells, "genes": n_genes, "regulons": R, "aucell_seconds": t1 - t0,
                      "cells_per_s": n_cells / (t1 - t0), "prepare_regulons_s": tb - ta, "dense_values_s": tc - tb,
                      "kernel_dtype": str(v.dtype)}), flush=True)
