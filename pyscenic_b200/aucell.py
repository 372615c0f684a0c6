# -*- coding: utf-8 -*-
"""
Drop-in replacement of ``pyscenic.aucell`` (reference ``src/pyscenic/aucell.py``) running on B200 GPUs.

Same public names, argument meaning, return types and error behaviour as the reference module:

    create_rankings(ex_mtx, seed=None)                                   aucell.py:25-51
    derive_auc_threshold(ex_mtx)                                         aucell.py:54-72
    enrichment(rnk_mtx, regulon, auc_threshold=0.05)                     aucell.py:75  (= enrichment4cells)
    aucell4r(df_rnk, signatures, auc_threshold, noweights, normalize, num_workers)     aucell.py:98-182
    aucell(exp_mtx, signatures, auc_threshold, noweights, normalize, seed, num_workers) aucell.py:185-213

The per-cell arithmetic (ranking with the seeded tie order, recovery-curve AUC) runs in hand-written sm_100a CUDA
kernels behind the C-ABI of ``libaucell_b200.so`` (``include/aucell_b200.h``).  There is no CPU fallback: without
the library or without a Blackwell GPU these functions raise.

Additive (keyword-only) extensions: ``exp_mtx`` may be a ``scipy.sparse`` matrix (CSR is consumed without
densifying) together with ``genes=`` / ``cells=`` labels; ``devices=`` selects the GPUs (cells are sharded
contiguously across them; the regulon tables are replicated).  ``num_workers`` is accepted for compatibility:
``num_workers == 1`` reproduces the reference's label-sorted output of its serial branch, any other value its
input-order output; it does not control parallelism (the GPUs do).
"""
from __future__ import annotations

import logging
import threading
from multiprocessing import cpu_count
from typing import List, Optional, Sequence

import numpy as np
import pandas as pd

from . import _native
from .plan import AUCellPlan, permutation_from_seed, prepare_regulons

LOGGER = logging.getLogger(__name__)
# The reference stores rankings as unsigned 32 bit integers (aucell.py:19-22).
DTYPE = "uint32"

__all__ = ["create_rankings", "derive_auc_threshold", "enrichment", "enrichment4cells", "aucell4r", "aucell"]


# ----------------------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------------------
def _default_devices(devices) -> List[int]:
    if devices is None:
        return [0]
    if isinstance(devices, int):
        return [devices]
    devs = [int(d) for d in devices]
    if not devs:
        raise ValueError("devices must name at least one GPU")
    return devs


def _is_sparse(x) -> bool:
    try:
        import scipy.sparse as sp

        return sp.issparse(x)
    except ImportError:  # pragma: no cover
        return False


def _dense_values(values: np.ndarray) -> np.ndarray:
    """Pick the narrowest float type that represents every value exactly: casting float64 -> float32 may create
    ties that do not exist in the input and would change the ranking."""
    v = np.asarray(values)
    if v.dtype == np.float32:
        return np.ascontiguousarray(v)
    if v.dtype == np.bool_:
        return np.ascontiguousarray(v, dtype=np.float32)
    v64 = np.ascontiguousarray(v, dtype=np.float64)
    if v.dtype.kind in "iu" and v.dtype.itemsize > 4 and v64.size:
        # int64 beyond 2^53 is not representable; the reference would rank the integers themselves
        if np.abs(v).max() > (1 << 53):
            raise ValueError("integer expression values beyond 2^53 are not supported")
    # cast + exactness check on all host cores (numpy: three single-threaded passes)
    v32 = np.empty(v64.shape, dtype=np.float32)
    exact = _native.load().aucell_cast_f64_to_f32(v64.ctypes.data, v32.ctypes.data, v64.size)
    if exact < 0:
        _native.check(exact)
    return v32 if exact == 1 else v64


def _torch_device(device: int):
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("pyscenic_b200 needs a CUDA device (sm_100); there is no CPU fallback")
    return torch.device("cuda", device)


def _shards(n_cells: int, n: int):
    bounds = [(n_cells * i) // n for i in range(n + 1)]
    return [(bounds[i], bounds[i + 1]) for i in range(n) if bounds[i + 1] > bounds[i]]


def _run_sharded(devices: List[int], n_cells: int, work) -> None:
    """Run ``work(device, c0, c1)`` for contiguous cell shards, one host thread per GPU (ctypes drops the GIL)."""
    shards = _shards(n_cells, len(devices))
    if len(shards) <= 1:
        for (c0, c1) in shards:
            work(devices[0], c0, c1)
        return
    errors = []

    def run(dev, c0, c1):
        try:
            work(dev, c0, c1)
        except BaseException as e:  # noqa: BLE001 - re-raised below
            errors.append(e)

    threads = [threading.Thread(target=run, args=(devices[i], c0, c1)) for i, (c0, c1) in enumerate(shards)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]


def _from_numpy(a: np.ndarray):
    """torch view of a host array that is only ever read (pandas hands out read-only arrays under copy-on-write)."""
    import warnings

    import torch

    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)
        return torch.from_numpy(a)


def _device_chunks(values: np.ndarray, device: int, chunk_bytes: int = 256 << 20):
    """Yield ``(c0, c1, tensor)``: consecutive cell ranges of a host matrix as CUDA tensors ``[cells x genes]`` in the
    narrowest exact float type, without ever making a transposing copy on the host.

    DataFrames usually hold their block gene-major (``to_numpy()`` is then a transposed view), and ``pd.read_csv`` gives
    float64.  numpy would need a strided 2-byte-at-a-time copy for the transpose and three full passes for the exactness
    check; here the host only copies contiguous row blocks in whatever order the memory already has, and the GPU does
    the transpose, the float64 -> float32 exactness test (casting may create ties that do not exist in the input, SURVEY
    7.3-f) and the cast."""
    import torch

    dev = _torch_device(device)
    v = np.asarray(values)
    if v.dtype == np.bool_ or (v.dtype.kind in "iu" and v.dtype.itemsize <= 2):
        v = v.astype(np.float32)
    elif v.dtype.kind in "iu":
        if v.size and v.dtype.itemsize > 4 and np.abs(v).max() > (1 << 53):
            raise ValueError("integer expression values beyond 2^53 are not supported")
        v = v.astype(np.float64)
    elif v.dtype not in (np.float32, np.float64):
        v = v.astype(np.float64)
    n_cells, n_genes = v.shape
    gene_major = (not v.flags.c_contiguous) and v.T.flags.c_contiguous
    chunk = max(1, chunk_bytes // (v.dtype.itemsize * max(n_genes, 1)))

    def finish(t):
        if t.dtype == torch.float64:
            t32 = t.to(torch.float32)
            if bool(((t32.to(torch.float64) == t) | torch.isnan(t)).all().item()):
                return t32     # every value of these cells survives the cast: same ranking, half the bytes
        return t

    if gene_major:
        vt = v.T                                      # [genes x cells], C-contiguous: the memory as it is
        free_bytes, _ = torch.cuda.mem_get_info(dev)
        if vt.nbytes <= free_bytes // 4:
            # upload the block as it lies in memory (contiguous gene rows, no host copy), transpose slices on the GPU
            tdtype = torch.float32 if v.dtype == np.float32 else torch.float64
            full = torch.empty((n_genes, n_cells), dtype=tdtype, device=dev)
            rows = max(1, chunk_bytes // (v.dtype.itemsize * max(n_cells, 1)))
            for g0 in range(0, n_genes, rows):
                full[g0 : g0 + rows].copy_(_from_numpy(vt[g0 : g0 + rows]), non_blocking=True)
            for c0 in range(0, n_cells, chunk):
                c1 = min(n_cells, c0 + chunk)
                yield c0, c1, finish(full[:, c0:c1].t().contiguous())
            return
        for c0 in range(0, n_cells, chunk):
            c1 = min(n_cells, c0 + chunk)
            # rows of the gene-major block are contiguous runs of cells: a plain block copy, then transpose on the GPU
            yield c0, c1, finish(_from_numpy(np.ascontiguousarray(vt[:, c0:c1])).to(dev).t().contiguous())
        return
    for c0 in range(0, n_cells, chunk):
        c1 = min(n_cells, c0 + chunk)
        yield c0, c1, finish(_from_numpy(np.ascontiguousarray(v[c0:c1])).to(dev))


def _dense_auc_shard(plan: AUCellPlan, values: np.ndarray, out: np.ndarray, device: int) -> None:
    # expression matrices are mostly zeros: the library squeezes them out on host threads, whatever the layout (a
    # DataFrame block is gene-major) and float type, and ranks the rest through the CSR kernels
    if plan.aucell_dense_strided_host(values, out):
        return
    if values.dtype == np.float32 and values.flags.c_contiguous:
        out[:] = plan.aucell_dense_host(values)       # library streams the host buffer itself
        return
    for c0, c1, t in _device_chunks(values, device):
        out[c0:c1] = plan.aucell_dense(t).cpu().numpy()


def _csr_auc_shard(plan: AUCellPlan, indptr, indices, data, out: np.ndarray, device: int) -> None:
    if data.dtype == np.float32:
        plan.aucell_csr_host(indptr, indices, data, out=out)
        return
    import torch

    dev = _torch_device(device)
    n_cells = len(indptr) - 1
    per = max(1.0, float(indptr[-1] - indptr[0]) / max(n_cells, 1))
    chunk = max(1, int((256 << 20) / (12.0 * per)))
    for c0 in range(0, n_cells, chunk):
        c1 = min(n_cells, c0 + chunk)
        e0, e1 = int(indptr[c0]), int(indptr[c1])
        ip = torch.from_numpy(np.ascontiguousarray(indptr[c0 : c1 + 1] - e0, dtype=np.int64)).to(dev)
        ix = torch.from_numpy(np.ascontiguousarray(indices[e0:e1], dtype=np.int32)).to(dev)
        dv = torch.from_numpy(np.ascontiguousarray(data[e0:e1])).to(dev)
        out[c0:c1] = plan.aucell_csr(ip, ix, dv).cpu().numpy()


def _indices_i32(indices: np.ndarray) -> np.ndarray:
    """Column indices as int32.  scipy switches to int64 indices beyond 2^31 stored entries (the 1.3 M-cell matrix):
    narrowed on all host cores by the library (numpy: one single-threaded pass over 20 GB)."""
    ix = np.asarray(indices)
    if ix.dtype == np.int32:
        return np.ascontiguousarray(ix)
    if ix.dtype != np.int64:
        return np.ascontiguousarray(ix, dtype=np.int32)
    ix = np.ascontiguousarray(ix)
    out = np.empty(ix.shape, dtype=np.int32)
    ok = _native.load().aucell_cast_i64_to_i32(ix.ctypes.data, out.ctypes.data, ix.size)
    if ok < 0:
        _native.check(ok)
    if ok == 0:
        raise ValueError("column indices outside [0, 2^31)")
    return out


def _canonical_csr(mtx):
    import scipy.sparse as sp

    csr = mtx if sp.isspmatrix_csr(mtx) else sp.csr_matrix(mtx)
    indptr = np.ascontiguousarray(csr.indptr, dtype=np.int64)
    indices = _indices_i32(csr.indices)
    # sorted, duplicate-free rows are what the kernels need; the library checks that on all host cores (scipy's own
    # has_canonical_format is a single-threaded scan: 2.5 s for the 1.3 M-cell benchmark matrix)
    ok = _native.load().aucell_csr_is_canonical(indptr.ctypes.data, indices.ctypes.data, csr.shape[0], csr.shape[1])
    if ok < 0:
        _native.check(ok)
    if ok == 0:
        csr = csr.copy()
        csr.sum_duplicates()
        indptr = np.ascontiguousarray(csr.indptr, dtype=np.int64)
        indices = _indices_i32(csr.indices)
    data = csr.data
    if data.dtype != np.float32:
        d = _dense_values(data)
        data = d
    return indptr, indices, data


def _assemble(values: np.ndarray, cells, names: Sequence[str], normalize: bool, num_workers: int) -> pd.DataFrame:
    aucs = pd.DataFrame(                       # (values is this call's own array: no defensive copy -- 0.7 s per 400 MB)
        data=values, index=pd.Index(data=cells, name="Cell"), columns=pd.Index(data=list(names), name="Regulon"), copy=False
    )
    if num_workers == 1:
        # serial branch of the reference (aucell.py:118-130): concat + unstack sorts both axes by label and
        # cannot reshape duplicate (Cell, Regulon) pairs
        if aucs.index.has_duplicates or aucs.columns.has_duplicates:
            raise ValueError("Index contains duplicate entries, cannot reshape")
        aucs = aucs.sort_index(axis=0).sort_index(axis=1)
    return aucs / aucs.max(axis=0) if normalize else aucs


# ----------------------------------------------------------------------------------------------------------
# public API
# ----------------------------------------------------------------------------------------------------------
def create_rankings(ex_mtx: pd.DataFrame, seed=None, *, device: int = 0) -> pd.DataFrame:
    """
    Create a whole genome rankings dataframe from a single cell expression profile dataframe
    (reference ``aucell.py:25-51``).

    :param ex_mtx: The expression profile matrix (n_cells x n_genes).
    :param seed: seed of the column shuffle that breaks ties (``None`` draws from the global numpy state).
    :return: uint32 rankings (n_cells x n_genes), 0 = most expressed; the columns are in SHUFFLED order, exactly as
        the reference returns them.
    """
    n_cells, n_genes = ex_mtx.shape
    perm = permutation_from_seed(n_genes, seed)
    out = np.empty((n_cells, n_genes), dtype=np.uint32)
    with AUCellPlan(n_genes, perm, None, device=device) as plan:
        for c0, c1, t in _device_chunks(ex_mtx.to_numpy(), device):
            out[c0:c1] = plan.rank_dense(t).cpu().numpy().view(np.uint32)
    return pd.DataFrame(data=out, index=ex_mtx.index, columns=ex_mtx.columns[perm])


def derive_auc_threshold(ex_mtx, *, device: int = 0) -> pd.Series:
    """
    Derive AUC thresholds for an expression matrix (reference ``aucell.py:54-72``): quantiles over cells of the
    fraction of detected (non-zero) genes.

    :param ex_mtx: The expression profile matrix (n_cells x n_genes): a DataFrame / array as in the reference, or a
        ``scipy.sparse`` matrix -- its stored values are counted on the GPU (``aucell_count_nonzero_csr_f32``; an
        explicit zero does not count, NaN does, like ``np.count_nonzero``) without densifying.
    :return: A pandas Series with the AUC thresholds at the 1 %, 5 %, 10 %, 50 % and 100 % quantiles.
    """
    if _is_sparse(ex_mtx):
        import ctypes

        import torch

        csr = ex_mtx.tocsr()
        dev = _torch_device(device)
        lib = _native.load()
        n_cells = csr.shape[0]
        counts = np.zeros(n_cells, dtype=np.int32)
        if n_cells and csr.nnz:
            values = np.asarray(csr.data)
            if values.dtype != np.float32:
                values = (values != 0).astype(np.float32)           # only "is it zero" matters; keeps NaN non-zero
            d_ip = _from_numpy(np.ascontiguousarray(csr.indptr, dtype=np.int64)).to(dev)
            d_dv = _from_numpy(np.ascontiguousarray(values)).to(dev)
            d_out = torch.empty(n_cells, dtype=torch.int32, device=dev)
            st = torch.cuda.current_stream(dev).cuda_stream
            _native.check(lib.aucell_count_nonzero_csr_f32(device, d_ip.data_ptr(), d_dv.data_ptr(), n_cells,
                                                           d_out.data_ptr(), ctypes.c_void_p(st)))
            counts = d_out.cpu().numpy()
        return pd.Series(counts).quantile([0.01, 0.05, 0.10, 0.50, 1]) / csr.shape[1]
    return pd.Series(np.count_nonzero(ex_mtx, axis=1)).quantile([0.01, 0.05, 0.10, 0.50, 1]) / ex_mtx.shape[1]


def enrichment4cells(rnk_mtx: pd.DataFrame, regulon, auc_threshold: float = 0.05, *, device: int = 0) -> pd.DataFrame:
    """
    ``ctxcore.recovery.enrichment4cells`` (the reference re-exports it as ``pyscenic.aucell.enrichment``,
    ``aucell.py:75``): AUC of one regulon for every cell of a ranking matrix.

    :return: DataFrame with MultiIndex ``(Cell, Regulon)`` and one float64 column ``"AUC"``.
    """
    res = _aucell4r_values(rnk_mtx, [regulon], auc_threshold, False, [device])
    index = pd.MultiIndex.from_arrays(
        [rnk_mtx.index.values, np.repeat(regulon.name, rnk_mtx.shape[0])], names=["Cell", "Regulon"]
    )
    return pd.DataFrame(index=index, data={"AUC": res[:, 0]})


enrichment = enrichment4cells


def _aucell4r_values(df_rnk: pd.DataFrame, signatures, auc_threshold, noweights, devices) -> np.ndarray:
    import torch

    n_cells, n_genes = df_rnk.shape
    tables = prepare_regulons(df_rnk.columns, None, signatures, auc_threshold, noweights)
    out = np.zeros((n_cells, len(tables.names)), dtype=np.float64)
    if not tables.valid.any() or n_cells == 0:
        return out
    ranks = np.ascontiguousarray(df_rnk.to_numpy(), dtype=np.uint32)

    def work(device, c0, c1):
        dev = _torch_device(device)
        with AUCellPlan(n_genes, None, tables, device=device) as plan:
            chunk = max(1, (256 << 20) // (4 * n_genes))
            for s in range(c0, c1, chunk):
                e = min(c1, s + chunk)
                t = torch.from_numpy(ranks[s:e].view(np.int32)).to(dev)
                out[s:e] = plan.from_rankings(t).cpu().numpy()

    _run_sharded(devices, n_cells, work)
    return out


def aucell4r(
    df_rnk: pd.DataFrame,
    signatures: Sequence,
    auc_threshold: float = 0.05,
    noweights: bool = False,
    normalize: bool = False,
    num_workers: int = cpu_count(),
    *,
    devices=None,
) -> pd.DataFrame:
    """
    Calculate enrichment of gene signatures for single cells from a rankings dataframe (reference
    ``aucell.py:98-182``).

    :param df_rnk: The rank matrix (n_cells x n_genes), 0-based.
    :param signatures: The gene signatures or regulons.
    :param auc_threshold: The fraction of the ranked genome to take into account for the AUC.
    :param noweights: ignore the gene weights of the signatures.
    :param normalize: Normalize the AUC values to a maximum of 1.0 per regulon.
    :param num_workers: accepted for compatibility (see module docstring).
    :return: A dataframe with the AUCs (n_cells x n_modules).
    """
    signatures = list(signatures)
    if len(signatures) == 0:
        raise ValueError("No signatures: nothing to score (the reference fails here as well)")
    devs = _default_devices(devices)
    values = _aucell4r_values(df_rnk, signatures, auc_threshold, noweights, devs)
    return _assemble(values, df_rnk.index.values, [s.name for s in signatures], normalize, num_workers)


def aucell(
    exp_mtx,
    signatures: Sequence,
    auc_threshold: float = 0.05,
    noweights: bool = False,
    normalize: bool = False,
    seed=None,
    num_workers: int = cpu_count(),
    *,
    genes: Optional[Sequence] = None,
    cells: Optional[Sequence] = None,
    devices=None,
) -> pd.DataFrame:
    """
    Calculate enrichment of gene signatures for single cells (reference ``aucell.py:185-213``).

    :param exp_mtx: The expression matrix (n_cells x n_genes): a DataFrame, or a scipy.sparse matrix together
        with ``genes`` (column labels) and optionally ``cells`` (row labels).
    :param signatures: The gene signatures or regulons.
    :param auc_threshold: The fraction of the ranked genome to take into account for the AUC.
    :param noweights: ignore the gene weights of the signatures.
    :param normalize: Normalize the AUC values to a maximum of 1.0 per regulon.
    :param seed: seed of the column shuffle that breaks ranking ties.
    :param num_workers: accepted for compatibility (see module docstring).
    :return: A dataframe with the AUCs (n_cells x n_modules).
    """
    signatures = list(signatures)
    if len(signatures) == 0:
        raise ValueError("No signatures: nothing to score (the reference fails here as well)")
    devs = _default_devices(devices)
    sparse = _is_sparse(exp_mtx)
    if sparse:
        if genes is None:
            raise ValueError("a sparse expression matrix needs genes= (column labels)")
        n_cells, n_genes = exp_mtx.shape
        gene_labels = np.asarray(genes, dtype=object)
        cell_labels = np.arange(n_cells) if cells is None else np.asarray(cells, dtype=object)
        if len(gene_labels) != n_genes or len(cell_labels) != n_cells:
            raise ValueError("genes/cells labels do not match the matrix shape")
    else:
        n_cells, n_genes = exp_mtx.shape
        gene_labels = exp_mtx.columns
        cell_labels = exp_mtx.index.values

    # the shuffle is drawn before anything else, like create_rankings() inside the reference's aucell()
    perm = permutation_from_seed(n_genes, seed)
    tables = prepare_regulons(gene_labels, perm, signatures, auc_threshold, noweights)
    out = np.zeros((n_cells, len(tables.names)), dtype=np.float64)

    if tables.valid.any() and n_cells > 0:
        if sparse:
            indptr, indices, data = _canonical_csr(exp_mtx)

            def work(device, c0, c1):
                with AUCellPlan(n_genes, perm, tables, device=device) as plan:
                    _csr_auc_shard(plan, indptr[c0 : c1 + 1], indices, data, out[c0:c1], device)

        else:
            values = exp_mtx.to_numpy()               # a view; dtype / layout are dealt with chunk by chunk on the GPU

            def work(device, c0, c1):
                with AUCellPlan(n_genes, perm, tables, device=device) as plan:
                    _dense_auc_shard(plan, values[c0:c1], out[c0:c1], device)

        _run_sharded(devs, n_cells, work)
    else:
        _native.load()  # fail loudly even when there is nothing to compute
    return _assemble(out, cell_labels, [s.name for s in signatures], normalize, num_workers)
