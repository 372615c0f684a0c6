#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""
Golden vectors for regulon specificity scores, made by EXECUTING THE REFERENCE's own module
(/root/reference/src/pyscenic/rss.py, which needs only numpy / pandas / scipy) in the build container => PINNED.

    python tests/golden/make_golden_rss.py        # needs /root/reference
"""
import importlib.util
import os

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    spec = importlib.util.spec_from_file_location("ref_rss", "/root/reference/src/pyscenic/rss.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rs = np.random.RandomState(12)
    n_cells, n_reg, n_types = 1500, 37, 7
    labels = rs.choice(["type%d" % t for t in range(n_types)], size=n_cells, p=[0.3, 0.25, 0.2, 0.1, 0.08, 0.05, 0.02])
    auc = rs.gamma(2.0, 0.02, size=(n_cells, n_reg))
    for r in range(n_reg):                       # regulons specific to a type, sparse regulons, one all-zero regulon
        t = r % n_types
        auc[labels == "type%d" % t, r] *= 1.0 + 0.5 * r
        if r % 5 == 0:
            auc[rs.rand(n_cells) < 0.6, r] = 0.0
    auc[:, 36] = 0.0
    cells = ["c%d" % i for i in range(n_cells)]
    auc_mtx = pd.DataFrame(auc, index=cells, columns=["R%d(+)" % r for r in range(n_reg)])
    cts = pd.Series(labels, index=cells)
    with np.errstate(all="ignore"):
        rss = ref.regulon_specificity_scores(auc_mtx, cts)
    assert rss.to_numpy().dtype == np.float32
    np.savez_compressed(os.path.join(HERE, "rss_golden.npz"), auc=auc, labels=labels.astype(str),
                        rss=rss.to_numpy(), types=np.asarray(rss.index, dtype=str))
    print("wrote rss_golden.npz", rss.shape, "NaN columns:", int(np.isnan(rss.to_numpy()).all(axis=0).sum()))


if __name__ == "__main__":
    main()
