# -*- coding: utf-8 -*-
"""Regulon specificity scores against outputs of the reference's own rss.py
(tests/golden/rss_golden.npz, made by tests/golden/make_golden_rss.py: PINNED)."""
import os

import numpy as np
import pandas as pd
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rss_golden.npz")


def _numpy_rss(auc, codes, n_types):
    """The closed form csrc/aucell_rss.cu evaluates, in numpy (host check of the algebra against the golden)."""
    out = np.empty((n_types, auc.shape[1]))
    with np.errstate(all="ignore"):
        p = auc / auc.sum(axis=0)
        for t in range(n_types):
            sel = codes == t
            u = 1.0 / sel.sum()
            pt = p[sel]
            s = u * np.log(2 * u / (pt + u)) + np.where(pt != 0, pt * np.log(2 * pt / (pt + u)), 0.0)
            out[t] = 1.0 - np.sqrt(0.5 * (np.log(2.0) * (1.0 - pt.sum(axis=0)) + s.sum(axis=0)))
    return out


def test_closed_form_matches_reference_golden():
    z = np.load(GOLD)
    types = list(z["types"])
    codes = np.asarray([types.index(x) for x in z["labels"]])
    got = _numpy_rss(z["auc"], codes, len(types)).astype(np.float32)
    exp = z["rss"]
    assert np.isnan(exp[:, 36]).all() and np.isnan(got[:, 36]).all()
    np.testing.assert_allclose(got[:, :36], exp[:, :36], rtol=2e-6, atol=2e-7)


@pytest.mark.gpu
def test_rss_gpu_matches_reference_golden():
    from pyscenic_b200.rss import regulon_specificity_scores

    z = np.load(GOLD)
    cells = ["c%d" % i for i in range(z["auc"].shape[0])]
    auc_mtx = pd.DataFrame(z["auc"], index=cells, columns=["R%d(+)" % r for r in range(z["auc"].shape[1])])
    got = regulon_specificity_scores(auc_mtx, pd.Series(z["labels"], index=cells))
    assert list(got.index) == list(z["types"]) and list(got.columns) == list(auc_mtx.columns)
    assert got.to_numpy().dtype == np.float32
    exp = z["rss"]
    assert np.isnan(got.to_numpy()[:, 36]).all()
    np.testing.assert_allclose(got.to_numpy()[:, :36], exp[:, :36], rtol=2e-6, atol=2e-7)
    # deterministic: two runs agree bit for bit
    again = regulon_specificity_scores(auc_mtx, pd.Series(z["labels"], index=cells))
    np.testing.assert_array_equal(got.to_numpy()[:, :36], again.to_numpy()[:, :36])


@pytest.mark.gpu
def test_rss_gpu_large_matches_closed_form():
    """200k cells x 600 regulons x 40 types: the GPU result against the numpy closed form (float64, 1e-9)."""
    import ctypes
    import torch
    from pyscenic_b200 import _native

    rs = np.random.RandomState(4)
    n_cells, n_reg, n_types = 200_000, 600, 40
    auc = rs.gamma(2.0, 0.02, size=(n_cells, n_reg))
    auc[rs.rand(n_cells, n_reg) < 0.3] = 0.0
    codes = rs.randint(0, n_types, size=n_cells).astype(np.int32)
    counts = np.bincount(codes, minlength=n_types).astype(np.int64)
    lib = _native.load()
    a = torch.from_numpy(auc).cuda()
    out = torch.empty((n_types, n_reg), dtype=torch.float64, device="cuda")
    _native.check(lib.aucell_rss_f64(0, a.data_ptr(), n_cells, n_reg, n_reg, torch.from_numpy(codes).cuda().data_ptr(),
                                     n_types, counts.ctypes.data_as(ctypes.c_void_p), out.data_ptr(), None))
    torch.cuda.synchronize()
    np.testing.assert_allclose(out.cpu().numpy(), _numpy_rss(auc, codes, n_types), rtol=1e-9, atol=1e-12)
    # error path: counts that do not add up
    bad = counts.copy(); bad[0] += 1
    rc = lib.aucell_rss_f64(0, a.data_ptr(), n_cells, n_reg, n_reg, torch.from_numpy(codes).cuda().data_ptr(), n_types,
                            bad.ctypes.data_as(ctypes.c_void_p), out.data_ptr(), None)
    assert rc == -1
