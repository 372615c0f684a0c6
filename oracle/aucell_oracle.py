# -*- coding: utf-8 -*-
"""
CPU oracle for the AUCell hot path of aertslab/pySCENIC  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module.  Nothing under ``pyscenic_b200/`` imports it.

Pinning status
--------------
* Ranking (``create_rankings``): PINNED.  The reference delegates the arithmetic to pandas
  (``src/pyscenic/aucell.py:46-51``); ``create_rankings`` below issues the same three pandas calls with
  the same arguments, and the independent restatements (``create_rankings_np`` in numpy,
  ``oracle_create_rankings_*`` in C) are checked against it and against golden vectors produced by
  executing the reference's own ``create_rankings`` (``tests/golden/make_golden.py``).
* AUC (``derive_rank_cutoff`` / ``enrichment4cells`` / ``aucs`` / ``auc2d`` / ``weighted_auc1d``):
  **PARITY UNPINNED.**  The arithmetic lives in the third-party package ``ctxcore>=0.2.0``
  (reference ``requirements.txt:1``; imported at ``src/pyscenic/aucell.py:14-15``), which is neither
  vendored under ``/root/reference`` nor installed in this image, and the reference's tests assert no
  AUC value (``tests/test_aucell.py:38-54`` are smoke tests over a missing fixture).  The functions below
  restate the published ctxcore 0.2.0 algorithm and are anchored on
  hand-computed known-answer tests (``tests/test_oracle.py``) and on the reference's call sites
  (``src/pyscenic/aucell.py:78-95, 118-181``), whose driver code is executed unmodified around these
  functions when the golden vectors are generated.
"""
from __future__ import annotations

import ctypes
import logging
import os
from itertools import repeat
from math import ceil
from multiprocessing import Array, Process, cpu_count
from multiprocessing.sharedctypes import RawArray
from operator import attrgetter, mul
from typing import Optional, Sequence

import numpy as np
import pandas as pd

LOGGER = logging.getLogger(__name__)
DTYPE = "uint32"  # reference src/pyscenic/aucell.py:21

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_aucell.so")
_lib = None


def build_c_oracle(force: bool = False) -> str:
    """Compile ``aucell_oracle.c`` with gcc (recipe: oracle/Makefile)."""
    import subprocess

    src = os.path.join(_HERE, "aucell_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle_aucell.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def c_lib():
    """ctypes handle of the C restatement (built on demand)."""
    global _lib
    if _lib is None:
        build_c_oracle()
        lib = ctypes.CDLL(_LIB_PATH)
        i64, p = ctypes.c_int64, ctypes.c_void_p
        lib.oracle_create_rankings_f32.argtypes = [p, i64, i64, p, p]
        lib.oracle_create_rankings_f32.restype = ctypes.c_int
        lib.oracle_create_rankings_f64.argtypes = [p, i64, i64, p, p]
        lib.oracle_create_rankings_f64.restype = ctypes.c_int
        lib.oracle_weighted_auc1d.argtypes = [p, p, i64, i64, ctypes.c_double]
        lib.oracle_weighted_auc1d.restype = ctypes.c_double
        lib.oracle_auc2d.argtypes = [p, i64, i64, p, i64, p, i64, ctypes.c_double, p]
        lib.oracle_auc2d.restype = ctypes.c_int
        _lib = lib
    return _lib


# ----------------------------------------------------------------------------------------------------
# Ranking  (reference src/pyscenic/aucell.py:25-51)
# ----------------------------------------------------------------------------------------------------
def create_rankings(ex_mtx: pd.DataFrame, seed=None) -> pd.DataFrame:
    """Same pandas call chain, same arguments, as reference ``aucell.py:46-51``."""
    return (
        ex_mtx.sample(frac=1.0, replace=False, axis=1, random_state=seed)
        .rank(axis=1, ascending=False, method="first", na_option="bottom")
        .astype(DTYPE)
        - 1
    )


def permutation_from_seed(n_genes: int, seed) -> np.ndarray:
    """The column permutation ``DataFrame.sample(frac=1, axis=1, random_state=seed)`` draws
    (pandas ``core/sample.py``: ``random_state.choice(n, size=n, replace=False)``; SURVEY 8a edge case 2)."""
    rs = pd.core.common.random_state(seed)
    return np.asarray(rs.choice(n_genes, size=n_genes, replace=False), dtype=np.int64)


def create_rankings_np(values: np.ndarray, perm: np.ndarray) -> np.ndarray:
    """Independent numpy restatement: stable descending sort of every row laid out in shuffled column
    order, NaN last.  Returns uint32 ``[n_cells, n_genes]`` in SHUFFLED column order (column j of the
    result is original column ``perm[j]``), like the frame the reference returns."""
    v = np.asarray(values)[:, perm].astype(np.float64)
    n_cells, n_genes = v.shape
    out = np.empty((n_cells, n_genes), dtype=np.uint32)
    for c in range(n_cells):
        row = v[c]
        nan = np.isnan(row)
        # descending stable: sort ascending on the negated values (NaN forced last); -(-0.0) == 0.0
        keyed = np.where(nan, np.inf, -row)
        order = np.lexsort((np.arange(n_genes), keyed, nan))  # primary: nan flag, then value, then position
        out[c, order] = np.arange(n_genes, dtype=np.uint32)
    return out


def create_rankings_c(values: np.ndarray, perm: np.ndarray) -> np.ndarray:
    """C restatement (aucell_oracle.c) -- used where the numpy loop is too slow."""
    lib = c_lib()
    perm = np.ascontiguousarray(perm, dtype=np.int64)
    n_cells, n_genes = values.shape
    out = np.empty((n_cells, n_genes), dtype=np.uint32)
    if values.dtype == np.float32:
        v = np.ascontiguousarray(values)
        rc = lib.oracle_create_rankings_f32(v.ctypes.data, n_cells, n_genes, perm.ctypes.data, out.ctypes.data)
    else:
        v = np.ascontiguousarray(values, dtype=np.float64)
        rc = lib.oracle_create_rankings_f64(v.ctypes.data, n_cells, n_genes, perm.ctypes.data, out.ctypes.data)
    if rc != 0:
        raise MemoryError("oracle_create_rankings failed")
    return out


def derive_auc_threshold(ex_mtx: pd.DataFrame) -> pd.Series:
    """reference ``aucell.py:54-72``."""
    return pd.Series(np.count_nonzero(ex_mtx, axis=1)).quantile([0.01, 0.05, 0.10, 0.50, 1]) / ex_mtx.shape[1]


# ----------------------------------------------------------------------------------------------------
# ctxcore.recovery restatement  (PARITY UNPINNED -- see module docstring)
# ----------------------------------------------------------------------------------------------------
def derive_rank_cutoff(auc_threshold: float, total_genes: int, rank_threshold: Optional[int] = None) -> int:
    """ctxcore.recovery.derive_rank_cutoff: ``int(round(thr * n)) - 1`` with Python's banker's rounding."""
    if not rank_threshold:
        rank_threshold = total_genes - 1
    assert 0 < rank_threshold < total_genes, "Rank threshold must be an integer between 1 and {0:d}".format(
        total_genes
    )
    assert 0.0 < auc_threshold <= 1.0, "AUC threshold must be a fraction between 0.0 and 1.0"
    rank_cutoff = int(round(auc_threshold * total_genes))
    assert 0 < rank_cutoff <= rank_threshold, (
        "An AUC threshold of {0:f} corresponds to {1:d} top ranked genes/regions in the database. "
        "Please increase the rank threshold or decrease the AUC threshold.".format(auc_threshold, rank_cutoff)
    )
    # the rank threshold itself is not included in the AUC calculation (R-SCENIC compatibility)
    return rank_cutoff - 1


def weighted_auc1d(ranking: np.ndarray, weights: np.ndarray, rank_cutoff: int, max_auc: float) -> float:
    """ctxcore.recovery.weighted_auc1d in plain numpy + sequential Python sums (small cases)."""
    filter_idx = ranking < rank_cutoff
    x = ranking[filter_idx].astype(np.int64)
    y = np.asarray(weights, dtype=np.float64)[filter_idx]
    sort_idx = np.argsort(x, kind="stable")
    x = np.concatenate((x[sort_idx], np.full((1,), rank_cutoff, dtype=np.int64)))
    cum = 0.0
    ycum = np.empty(len(sort_idx), dtype=np.float64)
    for i, w in enumerate(y[sort_idx]):  # sequential cumsum, as numba lowers it
        cum = cum + float(w)
        ycum[i] = cum
    total = 0.0
    for d, yy in zip(np.diff(x), ycum):  # sequential sum, as numba lowers np.sum
        total = total + float(d) * float(yy)
    return total / max_auc


def weighted_auc1d_c(ranking: np.ndarray, weights: np.ndarray, rank_cutoff: int, max_auc: float) -> float:
    r = np.ascontiguousarray(ranking, dtype=np.uint32)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    return c_lib().oracle_weighted_auc1d(r.ctypes.data, w.ctypes.data, len(r), int(rank_cutoff), float(max_auc))


def auc2d(rankings: np.ndarray, weights: np.ndarray, rank_cutoff: int, max_auc: float, use_c: bool = True) -> np.ndarray:
    """ctxcore.recovery.auc2d: one ``weighted_auc1d`` per row."""
    n_rows = rankings.shape[0]
    out = np.empty((n_rows,), dtype=np.float64)
    if use_c:
        r = np.ascontiguousarray(rankings, dtype=np.uint32)
        w = np.ascontiguousarray(weights, dtype=np.float64)
        cols = np.arange(r.shape[1], dtype=np.int64)
        rc = c_lib().oracle_auc2d(
            r.ctypes.data, n_rows, r.shape[1], cols.ctypes.data, r.shape[1], w.ctypes.data, int(rank_cutoff),
            float(max_auc), out.ctypes.data,
        )
        if rc != 0:
            raise MemoryError("oracle_auc2d failed")
    else:
        for i in range(n_rows):
            out[i] = weighted_auc1d(rankings[i, :], weights, rank_cutoff, max_auc)
    return out


def aucs(rnk: pd.DataFrame, total_genes: int, weights: np.ndarray, auc_threshold: float, use_c: bool = True) -> np.ndarray:
    """ctxcore.recovery.aucs."""
    rank_cutoff = derive_rank_cutoff(auc_threshold, total_genes)
    rankings = rnk.values
    y_max = weights.sum()
    # "+ 1": deliberate R-compatibility quirk of the reference implementation (SURVEY 8a edge case 5)
    maxauc = float((rank_cutoff + 1) * y_max)
    assert maxauc > 0
    return auc2d(rankings, weights, rank_cutoff, maxauc, use_c=use_c)


def enrichment4cells(rnk_mtx: pd.DataFrame, regulon, auc_threshold: float = 0.05, use_c: bool = True) -> pd.DataFrame:
    """ctxcore.recovery.enrichment4cells."""
    total_genes = len(rnk_mtx.columns)
    index = pd.MultiIndex.from_tuples(list(zip(rnk_mtx.index.values, repeat(regulon.name))), names=["Cell", "Regulon"])
    rnk = rnk_mtx.iloc[:, rnk_mtx.columns.isin(regulon.genes)]
    if rnk.empty or (float(len(rnk.columns)) / float(len(regulon))) < 0.80:
        LOGGER.warning(
            "Less than 80% of the genes in {} are present in the expression matrix.".format(regulon.name)
        )
        return pd.DataFrame(index=index, data={"AUC": np.zeros(shape=(rnk_mtx.shape[0]), dtype=np.float64)})
    weights = np.asarray([regulon[gene] for gene in rnk.columns.values])
    return pd.DataFrame(index=index, data={"AUC": aucs(rnk, total_genes, weights, auc_threshold, use_c=use_c)})


# ----------------------------------------------------------------------------------------------------
# Driver  (reference src/pyscenic/aucell.py:78-213)
# ----------------------------------------------------------------------------------------------------
def _chunked(src, size):
    """boltons.iterutils.chunked for the sizes the driver uses (size 0 is an error there too)."""
    if size <= 0:
        raise ValueError("expected a positive integer size")
    src = list(src)
    return [src[i : i + size] for i in range(0, len(src), size)]


def _enrichment(shared_ro_memory_array, modules, genes, cells, auc_threshold, auc_mtx, offset):
    """reference ``aucell.py:78-95`` (child process body)."""
    df_rnk = pd.DataFrame(
        data=np.frombuffer(shared_ro_memory_array, dtype=DTYPE).reshape(len(cells), len(genes)),
        columns=genes,
        index=cells,
    )
    result_mtx = np.frombuffer(auc_mtx.get_obj(), dtype="d")
    inc = len(cells)
    for idx, module in enumerate(modules):
        result_mtx[offset + (idx * inc) : offset + ((idx + 1) * inc)] = enrichment4cells(
            df_rnk, module, auc_threshold
        ).values.ravel(order="C")


def aucell4r(df_rnk, signatures, auc_threshold=0.05, noweights=False, normalize=False, num_workers=cpu_count()):
    """reference ``aucell.py:98-182``: label-sorting serial branch and fork/RawArray parallel branch."""
    if num_workers == 1:
        aucs_ = pd.concat(
            [
                enrichment4cells(df_rnk, module.noweights() if noweights else module, auc_threshold=auc_threshold)
                for module in signatures
            ]
        ).unstack("Regulon")
        aucs_.columns = aucs_.columns.droplevel(0)
    else:
        genes = df_rnk.columns.values
        cells = df_rnk.index.values
        from ctypes import c_uint32

        shared_ro_memory_array = RawArray(c_uint32, mul(*df_rnk.shape))
        array = np.frombuffer(shared_ro_memory_array, dtype=DTYPE)
        array[:] = df_rnk.values.ravel(order="C")
        auc_mtx = Array("d", len(cells) * len(signatures))
        if noweights:
            signatures = list(map(lambda m: m.noweights(), signatures))
        chunk_size = ceil(float(len(signatures)) / num_workers)
        processes = [
            Process(
                target=_enrichment,
                args=(shared_ro_memory_array, chunk, genes, cells, auc_threshold, auc_mtx, (chunk_size * len(cells)) * idx),
            )
            for idx, chunk in enumerate(_chunked(signatures, chunk_size))
        ]
        for p in processes:
            p.start()
        for p in processes:
            p.join()
        aucs_ = pd.DataFrame(
            data=np.ctypeslib.as_array(auc_mtx.get_obj()).reshape(len(signatures), len(cells)),
            columns=pd.Index(data=cells, name="Cell"),
            index=pd.Index(data=list(map(attrgetter("name"), signatures)), name="Regulon"),
        ).T
    return aucs_ / aucs_.max(axis=0) if normalize else aucs_


def aucell(exp_mtx, signatures, auc_threshold=0.05, noweights=False, normalize=False, seed=None, num_workers=cpu_count()):
    """reference ``aucell.py:185-213``."""
    return aucell4r(create_rankings(exp_mtx, seed), signatures, auc_threshold, noweights, normalize, num_workers)


# ----------------------------------------------------------------------------------------------------
# Array-level convenience for parity tests (no DataFrames): closed form of the per-cell AUC in the
# reference's own summation order, from a rank matrix in SHUFFLED column order.
# ----------------------------------------------------------------------------------------------------
def auc_matrix_from_ranks(
    ranks_shuffled: np.ndarray,
    reg_cols: Sequence[np.ndarray],
    reg_weights: Sequence[np.ndarray],
    reg_valid: Sequence[bool],
    rank_cutoff: int,
) -> np.ndarray:
    """``out[c, r]`` = weighted_auc1d of regulon r in cell c.  ``reg_cols[r]`` are shuffled-column indices
    in ascending order (the order ``rnk_mtx.iloc[:, columns.isin(genes)]`` selects them), ``reg_weights[r]``
    the matching weights; invalid (<80 % overlap) regulons give zeros."""
    lib = c_lib()
    r = np.ascontiguousarray(ranks_shuffled, dtype=np.uint32)
    n_cells = r.shape[0]
    out = np.zeros((n_cells, len(reg_cols)), dtype=np.float64)
    col = np.empty((n_cells,), dtype=np.float64)
    for k, (cols, w, ok) in enumerate(zip(reg_cols, reg_weights, reg_valid)):
        if not ok:
            continue
        cols = np.ascontiguousarray(cols, dtype=np.int64)
        w = np.ascontiguousarray(w, dtype=np.float64)
        maxauc = float((rank_cutoff + 1) * w.sum())
        assert maxauc > 0
        rc = lib.oracle_auc2d(
            r.ctypes.data, n_cells, r.shape[1], cols.ctypes.data, len(cols), w.ctypes.data, int(rank_cutoff), maxauc,
            col.ctypes.data,
        )
        if rc != 0:
            raise MemoryError("oracle_auc2d failed")
        out[:, k] = col
    return out
