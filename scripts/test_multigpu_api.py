"""In-process multi-GPU: aucell(..., devices=[0, 1]) must equal the single-GPU result bit for bit (dense f32 / f64,
gene-major frames, CSR)."""
import sys
sys.path.insert(0, ".")
import numpy as np, pandas as pd, scipy.sparse as sp, torch
import synth
from pyscenic_b200.aucell import aucell

assert torch.cuda.device_count() >= 2, torch.cuda.device_count()
rs = np.random.RandomState(0)
n_cells, n_genes = 3001, 5000
X = synth.dense_counts(rs, n_cells, n_genes, 0.1)
sigs = synth.regulons_to_signatures(*synth.make_regulons(rs, n_genes, 40, 10, 300, weighted=True))
cols = ["g%d" % j for j in range(n_genes)]
for name, df in (("f32", pd.DataFrame(X, columns=cols)), ("f64", pd.DataFrame(X.astype(np.float64) + 1e-13 * (X > 0), columns=cols)),
                 ("f32-cmajor", pd.DataFrame(np.ascontiguousarray(X), columns=cols, copy=False))):
    a = aucell(df, sigs, seed=3, devices=[0])
    b = aucell(df, sigs, seed=3, devices=[0, 1])
    c = aucell(df, sigs, seed=3, devices=[1])
    assert np.array_equal(a.values, b.values) and np.array_equal(a.values, c.values), name
    print(name, "ok", float(a.values.max()))
csr = sp.csr_matrix(X)
a = aucell(csr, sigs, seed=3, genes=cols, devices=[0])
b = aucell(csr, sigs, seed=3, genes=cols, devices=[0, 1])
assert np.array_equal(a.values, b.values)
print("csr ok")
