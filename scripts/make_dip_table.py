# -*- coding: utf-8 -*-
"""Distribution of Hartigan's dip statistic under the null (uniform samples), computed with THIS repository's GPU
implementation of the statistic (aucell_binarize_dip_f64): quantiles of sqrt(n) * dip for a ladder of sample sizes.
Writes pyscenic_b200/dip_table.json, which pyscenic_b200.binarization.dip_pvalues interpolates (the `diptest` package
ships such a table; it is not installed here, so the table is regenerated rather than copied).

    python scripts/make_dip_table.py            # needs a GPU; ~1 minute
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from pyscenic_b200.binarization import dip_statistics

NS = [4, 5, 6, 7, 8, 9, 10, 15, 20, 30, 50, 100, 200, 500, 1000, 2000, 5000, 10000, 20000, 40000, 72000]
PROBS = [0.0, 0.01, 0.02, 0.05, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 0.95, 0.98, 0.99, 0.995, 0.998, 0.999,
         0.9995, 0.9998, 0.9999, 1.0]


def main():
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(20261020)
    rows = []
    for n in NS:
        reps = 100_000 if n <= 1000 else (40_000 if n <= 10_000 else 10_000)
        dips = []
        done = 0
        while done < reps:
            b = min(reps - done, max(256, int(2e8 // n)))
            x = torch.rand((b, n), generator=g, device=dev, dtype=torch.float64)
            x = torch.sort(x, dim=1).values.contiguous()
            dips.append(dip_statistics(x).cpu().numpy())
            done += b
        d = np.sqrt(n) * np.concatenate(dips)
        rows.append([float(v) for v in np.quantile(d, PROBS)])
        print(n, reps, "sqrt(n)*dip at 0.5 / 0.95 / 0.99:", rows[-1][8], rows[-1][13], rows[-1][15], flush=True)
    out = {"n": NS, "prob": PROBS, "sqrt_n_dip": rows,
           "how": "quantiles of sqrt(n) * dip over uniform samples, scripts/make_dip_table.py (seed 20261020)"}
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pyscenic_b200", "dip_table.json")
    if len(sys.argv) > 1:
        path = sys.argv[1]
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
