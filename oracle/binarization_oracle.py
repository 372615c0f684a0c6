# -*- coding: utf-8 -*-
"""
CPU restatement of pyscenic.binarization (reference src/pyscenic/binarization.py:12-98).  TEST INFRASTRUCTURE ONLY:
nothing under pyscenic_b200/ may import this module.

``derive_threshold`` follows the reference line by line with the same libraries it calls (sklearn's GaussianMixture,
scipy's gaussian_kde and minimize_scalar); the one dependency that is not installed here, ``diptest``, is replaced by
``dip_statistic`` -- a plain-Python port of Hartigan & Hartigan's algorithm AS 217 (the published algorithm the package
wraps) -- with the p-value taken from the caller.  Parity status: mean + 2 std, mixture and trough branches are pinned
against the reference executed with a stubbed ``diptest`` (tests/golden/make_golden_binarization.py); the dip decision
itself is "parity unpinned".
"""
import numpy as np
from scipy import stats
from scipy.optimize import minimize_scalar
from sklearn import mixture


def dip_statistic(x_sorted):
    """Hartigan's dip of an ascending 1-D array (AS 217; greatest convex minorant / least concave majorant)."""
    x = np.concatenate([[0.0], np.asarray(x_sorted, dtype=np.float64)])    # 1-based
    n = len(x) - 1
    if n < 2 or x[n] == x[1]:
        return 0.0 if n < 1 else 1.0 / (2.0 * n)
    low, high, dip = 1, n, 1.0
    mn = [0] * (n + 2)
    mj = [0] * (n + 2)
    mn[1] = 1
    for j in range(2, n + 1):
        mn[j] = j - 1
        while True:
            mnj = mn[j]
            mnmnj = mn[mnj]
            if mnj == 1 or (x[j] - x[mnj]) * (mnj - mnmnj) < (x[mnj] - x[mnmnj]) * (j - mnj):
                break
            mn[j] = mnmnj
    mj[n] = n
    for k in range(n - 1, 0, -1):
        mj[k] = k + 1
        while True:
            mjk = mj[k]
            mjmjk = mj[mjk]
            if mjk == n or (x[k] - x[mjk]) * (mjk - mjmjk) < (x[mjk] - x[mjmjk]) * (k - mjk):
                break
            mj[k] = mjmjk
    while True:
        gcm = [0, high]
        while gcm[-1] > low:
            gcm.append(mn[gcm[-1]])
        l_gcm = len(gcm) - 1
        ig, ix = l_gcm, l_gcm - 1
        lcm = [0, low]
        while lcm[-1] < high:
            lcm.append(mj[lcm[-1]])
        l_lcm = len(lcm) - 1
        ih, iv = l_lcm, 2
        d = 0.0
        if l_gcm != 2 or l_lcm != 2:
            while True:
                gcmix, lcmiv = gcm[ix], lcm[iv]
                if gcmix > lcmiv:
                    gcmi1 = gcm[ix + 1]
                    dx = (lcmiv - gcmi1 + 1) - (x[lcmiv] - x[gcmi1]) * (gcmix - gcmi1) / (x[gcmix] - x[gcmi1])
                    iv += 1
                    if dx >= d:
                        d, ig, ih = dx, ix + 1, iv - 1
                else:
                    lcmiv1 = lcm[iv - 1]
                    dx = (x[gcmix] - x[lcmiv1]) * (lcmiv - lcmiv1) / (x[lcmiv] - x[lcmiv1]) - (gcmix - lcmiv1 - 1)
                    ix -= 1
                    if dx >= d:
                        d, ig, ih = dx, ix + 1, iv
                ix = max(ix, 1)
                iv = min(iv, l_lcm)
                if gcm[ix] == lcm[iv]:
                    break
        else:
            d = 1.0
        if d < dip:
            break
        dip_l = 0.0
        for j in range(ig, l_gcm):
            max_t = 1.0
            jb, je = gcm[j + 1], gcm[j]
            if je - jb > 1 and x[je] != x[jb]:
                c = (je - jb) / (x[je] - x[jb])
                for jj in range(jb, je + 1):
                    max_t = max(max_t, (jj - jb + 1) - (x[jj] - x[jb]) * c)
            dip_l = max(dip_l, max_t)
        dip_u = 0.0
        for j in range(ih, l_lcm):
            max_t = 1.0
            jb, je = lcm[j], lcm[j + 1]
            if je - jb > 1 and x[je] != x[jb]:
                c = (je - jb) / (x[je] - x[jb])
                for jj in range(jb, je + 1):
                    max_t = max(max_t, (x[jj] - x[jb]) * c - (jj - jb - 1))
            dip_u = max(dip_u, max_t)
        dip = max(dip, dip_l, dip_u)
        if low == gcm[ig] and high == lcm[ih]:
            break
        low, high = gcm[ig], lcm[ih]
    return dip / (2.0 * n)


def derive_threshold(data, bimodal, seed=None):
    """binarization.py:58-74 with the bimodality decision supplied by the caller."""
    data = np.asarray(data, dtype=np.float64)
    if not bimodal:
        return data.mean() + 2.0 * data.std()
    gmm2 = mixture.GaussianMixture(n_components=2, covariance_type="full", random_state=seed).fit(data.reshape(-1, 1))
    return minimize_scalar(fun=stats.gaussian_kde(data), bounds=sorted(gmm2.means_), method="bounded").x[0]


def isbimodal_bic(data, seed=None):
    """binarization.py:47-56."""
    X = np.asarray(data, dtype=np.float64).reshape(-1, 1)
    gmm2 = mixture.GaussianMixture(n_components=2, covariance_type="full", random_state=seed).fit(X)
    gmm1 = mixture.GaussianMixture(n_components=1, covariance_type="full", random_state=seed).fit(X)
    return bool(gmm2.bic(X) <= gmm1.bic(X))
