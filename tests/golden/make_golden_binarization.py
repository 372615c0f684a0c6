# -*- coding: utf-8 -*-
"""
Golden vectors for the binarization row: thresholds produced by EXECUTING the reference's own
src/pyscenic/binarization.py (sklearn GaussianMixture, scipy gaussian_kde + minimize_scalar as installed here).

The reference imports ``diptest``, which is not installed; it is stubbed with a function that returns the p-value this
script dictates per column -- 0.0 for columns that are bimodal beyond any doubt (two well separated groups), 1.0 for
plainly unimodal ones -- so the fixture pins everything BUT the dip test: the mean + 2 std branch, the mixture fit, the
kernel-density trough, and ``binarize``'s comparison.  Run from the repository root:

    python tests/golden/make_golden_binarization.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference/src/pyscenic/binarization.py"

_PVAL = {"value": 1.0}
stub = types.ModuleType("diptest")
stub.diptest = lambda *a, **k: (0.0, _PVAL["value"])
sys.modules["diptest"] = stub
spec = importlib.util.spec_from_file_location("ref_binarization", REF)
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


def columns(rs, n):
    cols, bimodal = [], []
    # unimodal: gamma, beta, normal, nearly constant
    cols.append(rs.gamma(2.0, 0.02, n)); bimodal.append(False)
    cols.append(rs.beta(2.0, 8.0, n) * 0.3); bimodal.append(False)
    cols.append(np.abs(rs.normal(0.08, 0.01, n))); bimodal.append(False)
    cols.append(np.full(n, 0.05) + rs.uniform(0, 1e-6, n)); bimodal.append(False)
    z = rs.gamma(1.5, 0.03, n); z[rs.rand(n) < 0.4] = 0.0
    cols.append(z); bimodal.append(False)                       # zero-inflated, still treated as unimodal
    # bimodal: two separated groups in different proportions and spreads
    for frac, (m0, s0), (m1, s1) in ((0.5, (0.03, 0.008), (0.15, 0.015)), (0.8, (0.02, 0.006), (0.12, 0.02)),
                                     (0.1, (0.05, 0.01), (0.2, 0.02)), (0.65, (0.01, 0.004), (0.06, 0.008)),
                                     (0.3, (0.04, 0.012), (0.11, 0.012))):
        k = int(frac * n)
        c = np.concatenate([np.abs(rs.normal(m0, s0, k)), np.abs(rs.normal(m1, s1, n - k))])
        rs.shuffle(c)
        cols.append(c); bimodal.append(True)
    return np.stack(cols, axis=1), np.asarray(bimodal)


def main():
    rs = np.random.RandomState(2026)
    out = {}
    for tag, n in (("a", 600), ("b", 2500)):
        auc, bimodal = columns(rs, n)
        names = ["R%d(+)" % i for i in range(auc.shape[1])]
        df = pd.DataFrame(auc, index=["c%d" % i for i in range(n)], columns=names)
        thr = []
        for name, b in zip(names, bimodal):
            _PVAL["value"] = 0.0 if b else 1.0
            thr.append(float(ref.derive_threshold(df, name, seed=42)))
        thr = np.asarray(thr)
        out[tag + "_auc"] = auc
        out[tag + "_bimodal"] = bimodal
        out[tag + "_thresholds"] = thr
        out[tag + "_binarized"] = (df > pd.Series(index=names, data=thr)).astype(int).to_numpy()
        # the BIC decision of the reference on the same columns (no stub involved)
        out[tag + "_bic_bimodal"] = np.asarray([
            float(ref.derive_threshold(df, name, seed=42, method="bic")) != float(df[name].values.mean() + 2.0 * df[name].values.std())
            for name in names])
        print(tag, np.round(thr, 5), out[tag + "_bic_bimodal"].astype(int))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "binarization_golden.npz"), **out)


if __name__ == "__main__":
    main()
