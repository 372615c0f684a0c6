# -*- coding: utf-8 -*-
"""
Host-side preparation of one AUCell job and the Python handle of the native plan.

What the reference does per regulon inside ``ctxcore.recovery.enrichment4cells`` / ``aucs`` (third-party; call
sites ``src/pyscenic/aucell.py:95,122-126``) and what is hoisted here, once per call, on the host:

* the seeded column permutation ``DataFrame.sample(frac=1, axis=1, random_state=seed)`` draws
  (``src/pyscenic/aucell.py:47``);
* ``rnk_mtx.columns.isin(regulon.genes)`` -> gene columns of every regulon (duplicate column names all count);
* the 80 % rule -> ``valid`` flag + the reference's warning text;
* ``weights = [regulon[gene] for gene in rnk.columns]`` in shuffled column order, ``max_auc =
  float((rank_cutoff + 1) * weights.sum())``;
* ``derive_rank_cutoff`` with Python's banker's ``round``.

Everything per cell happens in ``libaucell_b200.so``.
"""
from __future__ import annotations

import ctypes
import logging
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _native

LOGGER = logging.getLogger("pyscenic_b200.aucell")

__all__ = ["derive_rank_cutoff", "permutation_from_seed", "RegulonTables", "prepare_regulons", "AUCellPlan"]


def derive_rank_cutoff(auc_threshold: float, total_genes: int, rank_threshold: Optional[int] = None) -> int:
    """``ctxcore.recovery.derive_rank_cutoff``: ``int(round(auc_threshold * total_genes)) - 1`` (half-to-even
    rounding), with the reference's assertions (``AssertionError`` on a threshold outside ``(0, 1]`` or a cutoff
    outside ``(0, total_genes - 1]``)."""
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
    return rank_cutoff - 1


def _random_state(seed):
    """pandas' ``common.random_state`` (what ``DataFrame.sample(random_state=seed)`` applies): ``None`` is the
    global ``np.random`` state, RandomState/Generator objects pass through, anything else seeds a fresh legacy
    ``RandomState``."""
    if seed is None:
        return np.random
    if isinstance(seed, (np.random.RandomState, np.random.Generator)):
        return seed
    return np.random.RandomState(seed)


def permutation_from_seed(n_genes: int, seed) -> np.ndarray:
    """``perm[j]`` = gene column placed at shuffled position ``j`` -- the draw ``ex_mtx.sample(frac=1.0,
    replace=False, axis=1, random_state=seed)`` makes (``src/pyscenic/aucell.py:47``)."""
    rs = _random_state(seed)
    return np.asarray(rs.choice(n_genes, size=n_genes, replace=False), dtype=np.int64)


@dataclass
class RegulonTables:
    """Regulon CSR over gene COLUMN indices of the expression matrix + per-regulon constants."""

    names: List[str]
    indptr: np.ndarray        # int64 [R+1]
    cols: np.ndarray          # int32 [nnz] gene columns (original order of the matrix)
    weights: Optional[np.ndarray]  # float64 [nnz] or None when every weight is 1.0
    max_auc: np.ndarray       # float64 [R]
    valid: np.ndarray         # uint8 [R]
    rank_cutoff: int
    n_genes: int


def prepare_regulons(columns, perm: Optional[np.ndarray], signatures: Sequence, auc_threshold: float,
                     noweights: bool = False) -> RegulonTables:
    """Resolve signatures against the matrix' gene columns.

    ``columns`` are the gene labels in the matrix' own column order; ``perm`` is the shuffle (``None`` = the
    columns are already in ranking order, the ``aucell4r`` case)."""
    labels = np.asarray(columns, dtype=object)
    n_genes = len(labels)
    if perm is None:
        perm = np.arange(n_genes, dtype=np.int64)
    perm = np.asarray(perm, dtype=np.int64)
    # gene label -> shuffled position(s) holding that label.  Labels are normally unique: one dict built at C speed;
    # duplicate column names (all of them are selected, like isin) take the list-valued dict.
    shuffled_labels = labels[perm]
    pos_of = dict(zip(shuffled_labels.tolist(), range(n_genes)))
    multi = None
    if len(pos_of) != n_genes:
        multi = {}
        for j, g in enumerate(shuffled_labels):
            multi.setdefault(g, []).append(j)

    names, indptr, cols_l, w_l, max_auc, valid = [], [0], [], [], [], []
    rank_cutoff = None
    # one pass over all (gene, weight) pairs of all signatures: label -> position at C speed, one array conversion
    sigs = [sig.noweights() if noweights else sig for sig in signatures]
    keys, vals, bounds = [], [], [0]
    for sig in sigs:
        g2w = getattr(sig, "gene2weight", None)
        if hasattr(g2w, "keys"):
            keys.extend(g2w.keys())
            vals.extend(g2w.values())
        else:
            gs = list(sig.genes)
            keys.extend(gs)
            vals.extend(sig[g] for g in gs)
        bounds.append(len(keys))
    all_w = np.asarray(vals, dtype=np.float64)
    if multi is None:
        all_pos = np.asarray(list(map(pos_of.get, keys, [-1] * len(keys))), dtype=np.int64)
    for k, sig in enumerate(sigs):
        names.append(sig.name)
        b0, b1 = bounds[k], bounds[k + 1]
        if multi is None:
            ps = all_pos[b0:b1]
            present = ps >= 0
            sel = ps[present]
            w = all_w[b0:b1][present]
        else:
            hit = [(j, wt) for g, wt in zip(keys[b0:b1], vals[b0:b1]) for j in multi.get(g, ())]
            sel = np.asarray([h[0] for h in hit], dtype=np.int64)
            w = np.asarray([h[1] for h in hit], dtype=np.float64)
        n_sel = len(sel)
        if n_sel == 0 or (float(n_sel) / float(len(sig))) < 0.80:
            LOGGER.warning(
                "Less than 80% of the genes in {} are present in the expression matrix.".format(sig.name)
            )
            valid.append(0)
            max_auc.append(0.0)
            indptr.append(indptr[-1])
            continue
        if rank_cutoff is None:
            rank_cutoff = derive_rank_cutoff(auc_threshold, n_genes)
        order = np.argsort(sel, kind="stable")             # column order of rnk_mtx.iloc[:, isin] (positions are unique)
        sel = sel[order]
        w = np.ascontiguousarray(w[order])
        y_max = w.sum()
        m = float((rank_cutoff + 1) * y_max)
        assert m > 0
        valid.append(1)
        max_auc.append(m)
        cols_l.append(perm[sel].astype(np.int32))
        w_l.append(w)
        indptr.append(indptr[-1] + n_sel)
    cols = np.concatenate(cols_l).astype(np.int32) if cols_l else np.zeros(0, dtype=np.int32)
    weights = np.concatenate(w_l) if w_l else np.zeros(0, dtype=np.float64)
    unweighted = bool(np.all(weights == 1.0))
    return RegulonTables(
        names=names,
        indptr=np.asarray(indptr, dtype=np.int64),
        cols=cols,
        weights=None if unweighted else weights,
        max_auc=np.asarray(max_auc, dtype=np.float64),
        valid=np.asarray(valid, dtype=np.uint8),
        rank_cutoff=0 if rank_cutoff is None else int(rank_cutoff),
        n_genes=n_genes,
    )


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


def _torch_ptr(t):
    return None if t is None else t.data_ptr()


class AUCellPlan:
    """Python handle of ``aucell_plan_t``: permutation + regulon tables + cutoff resident on one device."""

    def __init__(self, n_genes: int, perm: Optional[np.ndarray], tables: Optional[RegulonTables] = None,
                 device: int = 0):
        self._lib = _native.load()
        self._handle = ctypes.c_void_p(None)
        self.device = int(device)
        self.n_genes = int(n_genes)
        self.tables = tables
        self.n_regulons = 0 if tables is None else len(tables.names)
        perm32 = None if perm is None else np.ascontiguousarray(perm, dtype=np.int32)
        if tables is None or self.n_regulons == 0:
            rc = self._lib.aucell_plan_create(
                self.device, self.n_genes, _ptr(perm32), 0, None, None, None, None, None, 0,
                ctypes.byref(self._handle),
            )
        else:
            t = tables
            self._keep = (
                np.ascontiguousarray(t.indptr, dtype=np.int64),
                np.ascontiguousarray(t.cols, dtype=np.int32),
                None if t.weights is None else np.ascontiguousarray(t.weights, dtype=np.float64),
                np.ascontiguousarray(t.max_auc, dtype=np.float64),
                np.ascontiguousarray(t.valid, dtype=np.uint8),
            )
            ip, cl, wt, ma, va = self._keep
            rc = self._lib.aucell_plan_create(
                self.device, self.n_genes, _ptr(perm32), self.n_regulons, _ptr(ip), _ptr(cl), _ptr(wt), _ptr(ma),
                _ptr(va), int(t.rank_cutoff), ctypes.byref(self._handle),
            )
        _native.check(rc)

    # ------------------------------------------------------------------ lifetime
    def close(self) -> None:
        if getattr(self, "_handle", None) is not None and self._handle.value:
            self._lib.aucell_plan_destroy(self._handle)
            self._handle = ctypes.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def path_stats(self, reset: bool = True) -> dict:
        """Code paths the fused kernel took since the last reset (see aucell_plan_path_stats)."""
        buf = (ctypes.c_uint64 * 5)()
        _native.check(self._lib.aucell_plan_path_stats(self._handle, buf, 1 if reset else 0))
        keys = ("cells", "general_too_many_entries", "general_negatives", "general_crowded_bin", "cells_with_zeros")
        return dict(zip(keys, [int(x) for x in buf]))

    def set_fast_path(self, enable: bool) -> bool:
        """Route CSR float32 calls through the tie-class kernel (default) or send every cell to the general kernel.
        Returns whether the plan qualifies for the tie-class kernel at all (see aucell_plan_set_fast_path)."""
        avail = ctypes.c_int(0)
        _native.check(self._lib.aucell_plan_set_fast_path(self._handle, 1 if enable else 0, ctypes.byref(avail)))
        return bool(avail.value)

    def fast_stats(self, reset: bool = True) -> dict:
        """Cells per outcome of the tie-class kernel since the last reset (see aucell_plan_fast_stats)."""
        buf = (ctypes.c_uint64 * 8)()
        _native.check(self._lib.aucell_plan_fast_stats(self._handle, buf, 1 if reset else 0))
        keys = ("ranked", "handed_long_row", "handed_negative_or_nan", "handed_duplicate_column", "handed_mixed_bin",
                "handed_tail_overflow", "handed_skipped")
        return dict(zip(keys, [int(x) for x in buf][:7]))

    def last_host_bytes(self):
        """(host-to-device, device-to-host) bytes the last ``*_host`` call moved (see aucell_plan_last_host_bytes)."""
        buf = (ctypes.c_int64 * 2)()
        _native.check(self._lib.aucell_plan_last_host_bytes(self._handle, buf))
        return int(buf[0]), int(buf[1])

    # ------------------------------------------------------------------ host buffers (library streams chunks)
    def aucell_dense_host(self, values: np.ndarray, want_nnz: bool = False, chunk_cells: int = 0):
        v = np.ascontiguousarray(values, dtype=np.float32)
        n_cells = v.shape[0]
        out = np.empty((n_cells, self.n_regulons), dtype=np.float64)
        nnz = np.empty((n_cells,), dtype=np.int32) if want_nnz else None
        _native.check(
            self._lib.aucell_dense_f32_host(self._handle, _ptr(v), n_cells, v.shape[1], _ptr(out), _ptr(nnz), chunk_cells)
        )
        return (out, nnz) if want_nnz else out

    def aucell_dense_strided_host(self, values: np.ndarray, out: np.ndarray) -> bool:
        """A dense host matrix as a DataFrame holds it (float32 / float64, cell-major or gene-major) through the library's
        squeeze-the-zeros-out path (``aucell_dense_strided_host``).  False: not applicable (layout, density, values that
        are not exact in float32) -- nothing computed, use another path."""
        v = values
        if v.ndim != 2 or v.dtype not in (np.float32, np.float64) or v.size == 0 or not out.flags.c_contiguous:
            return False
        es = v.dtype.itemsize
        if v.strides[0] % es or v.strides[1] % es:
            return False
        s0, s1 = v.strides[0] // es, v.strides[1] // es
        n_cells, n_genes = v.shape
        if not ((s1 == 1 and s0 >= n_genes) or (s0 == 1 and s1 >= n_cells)):
            return False
        rc = self._lib.aucell_dense_strided_host(self._handle, v.ctypes.data, 1 if v.dtype == np.float64 else 0, n_cells,
                                                 s0, s1, out.ctypes.data, None)
        if rc == 1:
            return False
        _native.check(rc)
        return True

    def aucell_csr_host(self, indptr: np.ndarray, indices: np.ndarray, data: np.ndarray, want_nnz: bool = False,
                        chunk_cells: int = 0, out: Optional[np.ndarray] = None):
        ip = np.ascontiguousarray(indptr, dtype=np.int64)
        ix = np.ascontiguousarray(indices, dtype=np.int32)
        dv = np.ascontiguousarray(data, dtype=np.float32)
        n_cells = len(ip) - 1
        if out is None:
            out = np.empty((n_cells, self.n_regulons), dtype=np.float64)
        nnz = np.empty((n_cells,), dtype=np.int32) if want_nnz else None
        _native.check(
            self._lib.aucell_csr_f32_host(self._handle, _ptr(ip), _ptr(ix), _ptr(dv), n_cells, _ptr(out), _ptr(nnz), chunk_cells)
        )
        return (out, nnz) if want_nnz else out

    # ------------------------------------------------------------------ device buffers (torch CUDA tensors)
    def _stream_ptr(self, stream):
        import torch

        # the current stream OF THE PLAN'S DEVICE (a worker thread of a multi-GPU call has device 0 current)
        s = torch.cuda.current_stream(self.device) if stream is None else stream
        return ctypes.c_void_p(s.cuda_stream)

    def _dense_fn(self, t, prefix):
        import torch

        if t.dtype == torch.float32:
            return getattr(self._lib, prefix + "_dense_f32")
        if t.dtype == torch.float64:
            return getattr(self._lib, prefix + "_dense_f64")
        raise TypeError("expression tensor must be float32 or float64, got {}".format(t.dtype))

    def _csr_fn(self, t, prefix):
        import torch

        if t.dtype == torch.float32:
            return getattr(self._lib, prefix + "_csr_f32")
        if t.dtype == torch.float64:
            return getattr(self._lib, prefix + "_csr_f64")
        raise TypeError("CSR data tensor must be float32 or float64, got {}".format(t.dtype))

    def aucell_dense(self, expr, out=None, out_nnz=None, stream=None):
        """expr: CUDA tensor [n_cells, n_genes] (row stride may exceed n_genes). Returns float64 [n_cells, R]."""
        import torch

        assert expr.is_cuda and expr.dim() == 2 and expr.stride(1) == 1 and expr.shape[1] == self.n_genes
        if out is None:
            out = torch.empty((expr.shape[0], self.n_regulons), dtype=torch.float64, device=expr.device)
        fn = self._dense_fn(expr, "aucell")
        _native.check(
            fn(self._handle, expr.data_ptr(), expr.shape[0], expr.stride(0) if expr.shape[0] > 1 else self.n_genes,
               out.data_ptr(), _torch_ptr(out_nnz), self._stream_ptr(stream))
        )
        return out

    def aucell_csr(self, indptr, indices, data, out=None, out_nnz=None, stream=None):
        """indptr int64 [n_cells+1], indices int32 (or uint16 with f32 data) [nnz], data f32/f64 [nnz] CUDA tensors."""
        import torch

        assert indptr.is_cuda and indptr.dtype == torch.int64
        n_cells = indptr.numel() - 1
        if out is None:
            out = torch.empty((n_cells, self.n_regulons), dtype=torch.float64, device=indptr.device)
        if indices.dtype in (torch.uint16, torch.int16):        # 16-bit column indices (n_genes <= 65535, float32 data)
            assert data.dtype == torch.float32
            fn = self._lib.aucell_csr16_f32
        else:
            assert indices.dtype == torch.int32
            fn = self._csr_fn(data, "aucell")
        _native.check(
            fn(self._handle, indptr.data_ptr(), indices.data_ptr(), data.data_ptr(), n_cells, out.data_ptr(),
               _torch_ptr(out_nnz), self._stream_ptr(stream))
        )
        return out

    def aucell_csr_lut8(self, indptr, indices16, codes, lut, out=None, out_nnz=None, scratch=None, stream=None):
        """CSR with uint16 (or int32) column indices and one-byte value codes into ``lut`` (float32 CUDA tensor, up to 256
        values): 3 (or 5) bytes per stored entry (``aucell_csr16_lut8_f32`` / ``aucell_csr_lut8_f32``).  ``scratch``:
        float32 [nnz] work space (allocated if None)."""
        import torch

        assert indptr.is_cuda and indptr.dtype == torch.int64 and codes.dtype == torch.uint8
        assert indices16.dtype in (torch.uint16, torch.int16, torch.int32) and lut.dtype == torch.float32 and lut.numel() <= 256
        fn = self._lib.aucell_csr_lut8_f32 if indices16.dtype == torch.int32 else self._lib.aucell_csr16_lut8_f32
        n_cells = indptr.numel() - 1
        if out is None:
            out = torch.empty((n_cells, self.n_regulons), dtype=torch.float64, device=indptr.device)
        lut256 = torch.zeros(256, dtype=torch.float32, device=indptr.device)
        lut256[: lut.numel()] = lut
        if scratch is None:
            scratch = torch.empty(max(int(codes.numel()), 1), dtype=torch.float32, device=indptr.device)
        _native.check(
            fn(self._handle, indptr.data_ptr(), indices16.data_ptr(), codes.data_ptr(), lut256.data_ptr(),
               scratch.data_ptr(), n_cells, out.data_ptr(), _torch_ptr(out_nnz), self._stream_ptr(stream))
        )
        return out

    def rank_dense(self, expr, out=None, out_nnz=None, stream=None):
        """Full ranking rows, uint32 stored in an int32 tensor [n_cells, n_genes], shuffled column order."""
        import torch

        assert expr.is_cuda and expr.dim() == 2 and expr.stride(1) == 1 and expr.shape[1] == self.n_genes
        if out is None:
            out = torch.empty((expr.shape[0], self.n_genes), dtype=torch.int32, device=expr.device)
        fn = self._dense_fn(expr, "aucell_rank")
        _native.check(
            fn(self._handle, expr.data_ptr(), expr.shape[0], expr.stride(0) if expr.shape[0] > 1 else self.n_genes,
               out.data_ptr(), _torch_ptr(out_nnz), self._stream_ptr(stream))
        )
        return out

    def rank_csr(self, indptr, indices, data, out=None, out_nnz=None, stream=None):
        import torch

        assert indptr.is_cuda and indptr.dtype == torch.int64 and indices.dtype == torch.int32
        n_cells = indptr.numel() - 1
        if out is None:
            out = torch.empty((n_cells, self.n_genes), dtype=torch.int32, device=indptr.device)
        fn = self._csr_fn(data, "aucell_rank")
        _native.check(
            fn(self._handle, indptr.data_ptr(), indices.data_ptr(), data.data_ptr(), n_cells, out.data_ptr(),
               _torch_ptr(out_nnz), self._stream_ptr(stream))
        )
        return out

    def from_rankings(self, ranks, out=None, stream=None):
        """ranks: int32/uint32-bit CUDA tensor [n_cells, n_genes] of 0-based ranks."""
        import torch

        assert ranks.is_cuda and ranks.dim() == 2 and ranks.stride(1) == 1 and ranks.element_size() == 4
        if out is None:
            out = torch.empty((ranks.shape[0], self.n_regulons), dtype=torch.float64, device=ranks.device)
        _native.check(
            self._lib.aucell_from_rankings_u32(
                self._handle, ranks.data_ptr(), ranks.shape[0],
                ranks.stride(0) if ranks.shape[0] > 1 else self.n_genes, out.data_ptr(), self._stream_ptr(stream))
        )
        return out
