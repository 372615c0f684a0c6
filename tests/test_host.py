# -*- coding: utf-8 -*-
"""CPU tests of the host side: signature records, regulon resolution (isin / 80 % rule / weights / max_auc),
the seeded permutation, the rank cutoff, and that the C-ABI library loads and exports every declared symbol."""
import ctypes
import os
import re

import numpy as np
import pandas as pd
import pytest

from oracle import aucell_oracle as O
from pyscenic_b200 import _native
from pyscenic_b200.genesig import GeneSignature, Regulon
from pyscenic_b200.plan import derive_rank_cutoff, permutation_from_seed, prepare_regulons

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GMT = "/root/reference/src/resources/tests/c6.all.v6.1.symbols.gmt"


# ------------------------------------------------------------------------------------------ C-ABI library
def test_library_loads_and_exports_every_declared_symbol():
    lib = _native.load()
    header = open(os.path.join(ROOT, "include", "aucell_b200.h")).read()
    declared = set(re.findall(r"\b(aucell_[a-z0-9_]+)\s*\(", header))
    declared -= {"aucell_plan_t"}
    assert declared == set(_native.EXPORTED_SYMBOLS)
    for sym in declared:
        assert getattr(lib, sym) is not None
    want = int(re.search(r"#define\s+AUCELL_B200_ABI_VERSION\s+(\d+)", header).group(1))
    assert lib.aucell_b200_abi_version() == want
    assert lib.aucell_b200_launch_count() >= 0


def test_csr_is_canonical_host_check():
    """aucell_csr_is_canonical (host threads, no GPU): what scipy's has_canonical_format says, plus the column range."""
    import scipy.sparse as sp

    lib = _native.load()

    def check(indptr, indices, n_genes):
        ip = np.ascontiguousarray(indptr, dtype=np.int64)
        ix = np.ascontiguousarray(indices, dtype=np.int32)
        return lib.aucell_csr_is_canonical(ip.ctypes.data, ix.ctypes.data if len(ix) else None, len(ip) - 1, n_genes)

    rs = np.random.RandomState(0)
    m = sp.random(300, 500, density=0.05, format="csr", random_state=rs, dtype=np.float32)
    m.sort_indices()
    assert m.has_canonical_format and check(m.indptr, m.indices, 500) == 1
    assert check(m.indptr, m.indices, int(m.indices.max())) == 0                  # a column out of range
    bad = m.indices.copy()
    r = int(np.argmax(np.diff(m.indptr) >= 2))
    s0 = m.indptr[r]
    bad[s0], bad[s0 + 1] = bad[s0 + 1], bad[s0]                                  # one row unsorted
    assert check(m.indptr, bad, 500) == 0
    dup = m.indices.copy()
    dup[s0 + 1] = dup[s0]                                                        # a duplicate column
    assert check(m.indptr, dup, 500) == 0
    neg = m.indices.copy()
    neg[s0] = -1
    assert check(m.indptr, neg, 500) == 0
    assert check([0, 0, 0], [], 10) == 1                                         # empty rows
    assert check([0], [], 10) == 1                                               # no rows
    assert lib.aucell_csr_is_canonical(None, None, 3, 10) < 0                    # bad arguments
    big = sp.random(20000, 2000, density=0.02, format="csr", random_state=rs, dtype=np.float32)   # several threads
    big.sort_indices()
    assert check(big.indptr, big.indices, 2000) == 1
    ix = big.indices.copy()
    ix[-1] = 2000
    assert check(big.indptr, ix, 2000) == 0


def test_library_has_sm100a_code_only():
    import subprocess, shutil

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", _native.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pyscenic_b200.aucell import aucell

    df = pd.DataFrame(np.eye(4, 30, dtype=np.float32), columns=["g%d" % i for i in range(30)])
    sig = GeneSignature("s", ["g0", "g1", "g2"])
    with pytest.raises(_native.AUCellNativeError):
        aucell(df, [sig], seed=1)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under pyscenic_b200/ may import, load or mention it."""
    pkg = os.path.join(ROOT, "pyscenic_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower(), "%s mentions the oracle" % f


# ------------------------------------------------------------------------------------------ host arithmetic
def test_rank_cutoff_equals_oracle():
    for n in (100, 2000, 2010, 2030, 20000, 25000, 27998, 70001):
        for thr in (0.01, 0.02, 0.05, 0.1, 0.2, 0.5):
            assert derive_rank_cutoff(thr, n) == O.derive_rank_cutoff(thr, n)
    for bad in (0.0, 1.0001, -1):
        with pytest.raises(AssertionError):
            derive_rank_cutoff(bad, 100)


def test_permutation_is_what_pandas_sample_draws():
    for seed in (0, 42, 2**31 - 1):
        n = 257
        df = pd.DataFrame(np.zeros((1, n)), columns=np.arange(n))
        want = df.sample(frac=1.0, replace=False, axis=1, random_state=seed).columns.values
        assert np.array_equal(permutation_from_seed(n, seed), want)
    # RandomState objects pass through (and advance), None uses the global numpy state
    rs1, rs2 = np.random.RandomState(5), np.random.RandomState(5)
    df = pd.DataFrame(np.zeros((1, 50)), columns=np.arange(50))
    assert np.array_equal(permutation_from_seed(50, rs1), df.sample(frac=1.0, axis=1, random_state=rs2).columns.values)
    np.random.seed(77)
    a = permutation_from_seed(50, None)
    np.random.seed(77)
    b = df.sample(frac=1.0, axis=1, random_state=None).columns.values
    assert np.array_equal(a, b)


def test_prepare_regulons_matches_reference_membership_rules(caplog):
    genes = ["g%d" % i for i in range(20)] + ["g3"]          # duplicated column name: both columns selected
    perm = np.random.RandomState(1).permutation(len(genes))
    sigs = [
        GeneSignature("full", {"g0": 2.0, "g1": 0.5, "g3": 1.5}),
        GeneSignature("eighty", {"g0": 1.0, "g1": 1.0, "g2": 1.0, "g4": 1.0, "zz": 1.0}),      # 4/5 = 0.8 passes
        GeneSignature("below", {"g0": 1.0, "g1": 1.0, "g2": 1.0, "zz": 1.0, "yy": 1.0}),       # 3/5 < 0.8
        GeneSignature("none", ["qq", "rr"]),
    ]
    with caplog.at_level("WARNING"):
        t = prepare_regulons(genes, perm, sigs, 0.25, noweights=False)
    assert t.valid.tolist() == [1, 1, 0, 0]
    assert "Less than 80% of the genes in below are present in the expression matrix." in caplog.text
    assert "Less than 80% of the genes in none are present in the expression matrix." in caplog.text
    assert t.rank_cutoff == O.derive_rank_cutoff(0.25, 21)
    # "full": g0, g1 and BOTH g3 columns, in shuffled column order
    cols = t.cols[t.indptr[0]:t.indptr[1]]
    assert sorted(np.asarray(genes)[cols]) == ["g0", "g1", "g3", "g3"]
    inv = np.empty_like(perm); inv[perm] = np.arange(len(perm))
    assert np.all(np.diff(inv[cols]) > 0)
    w = t.weights[t.indptr[0]:t.indptr[1]]
    assert np.array_equal(w, [sigs[0][genes[c]] for c in cols])
    assert t.max_auc[0] == float((t.rank_cutoff + 1) * w.sum())
    # same answer as the oracle's frame-level enrichment4cells on a real ranking
    x = np.random.RandomState(2).poisson(1.0, size=(7, len(genes))).astype(np.float32)
    df = pd.DataFrame(x, columns=genes)
    rnk = O.create_rankings(df, seed=1)
    for k, s in enumerate(sigs):
        e = O.enrichment4cells(rnk, s, 0.25)["AUC"].values
        if not t.valid[k]:
            assert (e == 0).all()
    # noweights -> integer path (weights None)
    t2 = prepare_regulons(genes, perm, sigs, 0.25, noweights=True)
    assert t2.weights is None and t2.max_auc[0] == float((t2.rank_cutoff + 1) * 4)


def test_all_invalid_signatures_do_not_assert_on_threshold():
    # the reference only evaluates the cutoff inside aucs(), i.e. for regulons that pass the 80 % rule
    t = prepare_regulons(["a", "b", "c"], None, [GeneSignature("x", ["q"])], 5.0)
    assert t.valid.tolist() == [0]
    with pytest.raises(AssertionError):
        prepare_regulons(["a", "b", "c"], None, [GeneSignature("x", ["a"])], 5.0)


# ------------------------------------------------------------------------------------------ signatures
def test_gene_signature_record():
    s = GeneSignature("s", ["a", "b", "c"])
    assert len(s) == 3 and s["a"] == 1.0 and "b" in s and set(s.genes) == {"a", "b", "c"}
    w = GeneSignature("w", {"a": 0.1, "b": 3.0, "c": 2.0})
    assert w.genes == ("b", "c", "a") and w.weights == (3.0, 2.0, 0.1)
    assert w.noweights().weights == (1.0, 1.0, 1.0) and w.noweights().name == "w"
    assert GeneSignature("p", [("a", 2.0), ("b", 1.0)])["a"] == 2.0
    u = w.union(GeneSignature("v", {"a": 5.0, "d": 1.0}))
    assert u["a"] == 5.0 and len(u) == 4
    assert w.rename("z").name == "z" and w.add("q", 9.0)["q"] == 9.0
    assert len(w.difference(s)) == 0 and len(w.intersection(s)) == 3
    with pytest.raises(AttributeError):
        w.name = "other"
    with pytest.raises(ValueError):
        GeneSignature("", ["a"])
    assert hash(s) == hash(GeneSignature("s", ["a", "b", "c"])) and s == GeneSignature("s", ["c", "b", "a"])


def test_regulon_record():
    r = Regulon("TF1(+)", {"a": 1.0, "b": 2.0}, transcription_factor="TF1", context=frozenset(["activating"]))
    assert r.transcription_factor == "TF1" and "activating" in r.context
    assert isinstance(r.noweights(), Regulon) and r.noweights()["b"] == 1.0
    r2 = r.union(Regulon("TF1(+)", {"b": 5.0, "c": 1.0}, transcription_factor="TF1", score=3.0))
    assert r2["b"] == 5.0 and r2.score == 3.0 and len(r2) == 3
    with pytest.raises(ValueError):
        Regulon("x", ["a"], transcription_factor="")


def test_gmt_round_trip(tmp_path):
    sigs = [GeneSignature("A", ["g1", "g2", "g3"]), GeneSignature("B", ["g2"])]
    p = str(tmp_path / "x.gmt")
    GeneSignature.to_gmt(p, sigs, field_separator="\t", gene_separator="\t")
    back = GeneSignature.from_gmt(p, field_separator="\t", gene_separator="\t")
    assert [s.name for s in back] == ["A", "B"] and set(back[0].genes) == {"g1", "g2", "g3"}


@pytest.mark.skipif(not os.path.exists(GMT), reason="reference fixture only exists in the build container")
def test_reference_gmt_fixture_loads():
    # the signature fixture the reference's own tests use (tests/test_aucell.py:14,22-28)
    gs = GeneSignature.from_gmt(GMT, gene_separator="\t", field_separator="\t")
    assert len(gs) == 189
    assert min(len(s) for s in gs) == 10 and max(len(s) for s in gs) == 481
    assert sum(len(s) for s in gs) == 31319


def _build_worker(args):
    """One rank of a multi-process job asking for the library (fork: inherits the patched module state)."""
    from pyscenic_b200 import _native

    _native.build()
    return os.getpid()


def test_concurrent_builds_are_serialised_and_decided_by_content(tmp_path, monkeypatch):
    """N ranks of one job must not all rebuild the library (seen on an 8-GPU box: every rank found the copied library
    'older' than a source file and raced on the link step).  Staleness is a content hash; builds take a file lock and
    late-comers find the work done."""
    import multiprocessing as mp
    import time

    from pyscenic_b200 import _native

    lib = str(tmp_path / "libfake.so")
    counter = tmp_path / "builds.txt"
    monkeypatch.setattr(_native, "LIB_PATH", lib)
    monkeypatch.setattr(_native, "_HASH_PATH", lib + ".src.sha256")
    monkeypatch.setattr(_native, "_LOCK_PATH", lib + ".lock")
    monkeypatch.setattr(_native, "_nvcc", lambda: "/bin/true")

    def fake_build(nvcc, verbose):
        time.sleep(0.3)
        with open(counter, "a") as fh:
            fh.write("built by %d\n" % os.getpid())
        with open(lib, "wb") as fh:
            fh.write(b"\x7fELF")
        with open(lib + ".src.sha256", "w") as fh:
            fh.write(_native._source_hash() + "\n")
        return lib

    monkeypatch.setattr(_native, "_build_locked", fake_build)
    assert _native.is_stale()
    ctx = mp.get_context("fork")
    with ctx.Pool(4) as pool:
        pids = pool.map(_build_worker, range(4))
    assert len(set(pids)) >= 2
    assert counter.read_text().count("built by") == 1
    assert not _native.is_stale()
    # a copy that scrambles modification times does not trigger a rebuild; a changed source does
    os.utime(lib, (1, 1))
    assert not _native.is_stale()
    with open(lib + ".src.sha256", "w") as fh:
        fh.write("0" * 64 + "\n")
    assert _native.is_stale()


def test_host_cast_helpers():
    """aucell_cast_f64_to_f32 / aucell_cast_i64_to_i32 (host threads, no GPU) behind aucell._dense_values / _indices_i32."""
    from pyscenic_b200.aucell import _dense_values, _indices_i32

    rs = np.random.RandomState(1)
    counts = np.floor(rs.exponential(3.0, size=3_000_000))
    v = _dense_values(counts)
    assert v.dtype == np.float32 and np.array_equal(v.astype(np.float64), counts)
    special = np.array([1.0, np.nan, np.inf, -np.inf, -0.0, 2.5, 2.0 ** -149])
    v = _dense_values(special)
    assert v.dtype == np.float32 and np.array_equal(v.astype(np.float64), special, equal_nan=True)
    assert np.signbit(v[4])
    for bad in (0.1, 1e300, 2.0 ** -150, 16777217.0):                  # not float32 values: the matrix stays float64
        x = counts.copy()
        x[-1] = bad
        assert _dense_values(x).dtype == np.float64
    assert _dense_values(np.arange(7, dtype=np.int64)).dtype == np.float32
    ix = rs.randint(0, 28000, size=2_000_000).astype(np.int64)
    got = _indices_i32(ix)
    assert got.dtype == np.int32 and np.array_equal(got, ix)
    for bad in (2 ** 31, -1):
        y = ix.copy()
        y[123] = bad
        with pytest.raises(ValueError):
            _indices_i32(y)
    assert _indices_i32(np.array([3, 4], dtype=np.int32)).dtype == np.int32


def _host_dense_to_csr(arr, n_cells=None, n_genes=None):
    """aucell_host_dense_to_csr on a 2-D numpy array in either layout; returns (rc, indptr, indices, values)."""
    import ctypes

    lib = _native.load()
    es = arr.dtype.itemsize
    nc, ng = arr.shape if n_cells is None else (n_cells, n_genes)
    s0, s1 = arr.strides[0] // es, arr.strides[1] // es
    indptr = np.full(nc + 1, -1, dtype=np.int64)
    cap = max(1, nc * ng)
    indices = np.full(cap, -1, dtype=np.int32)
    values = np.full(cap, -1.0, dtype=np.float32)
    nnz = ctypes.c_int64(-1)
    rc = lib.aucell_host_dense_to_csr(arr.ctypes.data, 1 if arr.dtype == np.float64 else 0, nc, ng, s0, s1,
                                      indptr.ctypes.data, indices.ctypes.data, values.ctypes.data, cap, ctypes.byref(nnz))
    return rc, indptr, indices[: max(nnz.value, 0)], values[: max(nnz.value, 0)], nnz.value


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("order", ["C", "F"])
def test_host_dense_to_csr_equals_scipy(dtype, order):
    """The squeeze behind aucell_dense_strided_host (host threads, AVX2; no GPU) against scipy, for the layouts and float
    types a DataFrame block comes in: ragged tile and vector tails, -0.0 / NaN / inf, an all-zero cell, one cell far denser
    than the rest (the per-cell capacity of a gene-major tile has to grow), padded strides."""
    import scipy.sparse as sp

    rs = np.random.RandomState(2)
    n_cells, n_genes = 1337, 2051
    X = np.where(rs.random_sample((n_cells, n_genes)) < 0.07, np.floor(1 + rs.exponential(2.0, (n_cells, n_genes))), 0.0)
    X[3] = 0.0
    X[5, :] = np.arange(1, n_genes + 1) % 7 + 1.0                  # a dense cell among sparse ones
    X[7, 11] = np.nan
    X[8, 12] = np.inf
    X[9, rs.choice(n_genes, 50, replace=False)] = -0.0
    X[10, -3:] = 2.0
    X[-1, -1] = 5.0
    pad = np.full((n_cells + 5, n_genes + 3), 9.0)                 # (strides larger than the matrix)
    pad[:n_cells, :n_genes] = X
    big = np.asarray(pad, dtype=dtype, order=order)
    arr = big[:n_cells, :n_genes]
    rc, indptr, indices, values, nnz = _host_dense_to_csr(arr)
    assert rc == 0
    want = sp.csr_matrix(np.where(np.isnan(X), 1.0, X))           # (the pattern; NaN is a stored value)
    want.sort_indices()
    assert nnz == want.nnz and np.array_equal(indptr, want.indptr) and np.array_equal(indices, want.indices)
    got_vals = np.asarray(X[np.repeat(np.arange(n_cells), np.diff(want.indptr)), want.indices], dtype=np.float32)
    assert np.array_equal(values, got_vals, equal_nan=True)
    assert not np.any(values == 0.0)                               # +0.0 / -0.0 are dropped


def test_host_dense_to_csr_declines_and_errors():
    rs = np.random.RandomState(4)
    X = np.where(rs.random_sample((300, 500)) < 0.1, 1.0, 0.0)
    for order in ("C", "F"):
        inexact = np.asarray(X, dtype=np.float64, order=order).copy(order=order)
        inexact[17, 33] = 0.1
        rc = _host_dense_to_csr(inexact)[0]
        assert rc == 1                                             # AUCELL_NOT_SPARSE: keep float64
    lib = _native.load()
    import ctypes

    a = np.asarray(X, dtype=np.float32)
    indptr = np.zeros(301, dtype=np.int64)
    nnz = ctypes.c_int64(0)
    small = np.zeros(3, dtype=np.int32)
    smallv = np.zeros(3, dtype=np.float32)
    rc = lib.aucell_host_dense_to_csr(a.ctypes.data, 0, 300, 500, 500, 1, indptr.ctypes.data, small.ctypes.data,
                                      smallv.ctypes.data, 3, ctypes.byref(nnz))
    assert rc == -5 and nnz.value == int(np.count_nonzero(X)) and indptr[-1] == nnz.value      # AUCELL_ERR_NOMEM, sizes known
    rc = lib.aucell_host_dense_to_csr(a.ctypes.data, 0, 300, 500, 7, 3, indptr.ctypes.data, small.ctypes.data,
                                      smallv.ctypes.data, 3, ctypes.byref(nnz))
    assert rc == -1                                                # neither stride is 1
    ip = np.full(1, -1, dtype=np.int64)
    rc = lib.aucell_host_dense_to_csr(a.ctypes.data, 0, 0, 10, 10, 1, ip.ctypes.data, None, None, 0, ctypes.byref(nnz))
    assert rc == 0 and nnz.value == 0 and ip[0] == 0              # no cells


def test_host_worker_pool_serves_concurrent_callers():
    """The persistent host workers (one parallel region at a time, regions of different callers queue): several Python
    threads call multi-threaded library functions at once, with results checked, and nothing deadlocks."""
    import threading

    lib = _native.load()
    rs = np.random.RandomState(3)
    src = [np.floor(rs.exponential(3.0, size=(1 << 21) + 17 * k)) for k in range(4)]
    errors = []

    def caller(k):
        try:
            for _ in range(6):
                dst = np.empty(src[k].shape, dtype=np.float32)
                assert lib.aucell_cast_f64_to_f32(src[k].ctypes.data, dst.ctypes.data, src[k].size) == 1
                assert np.array_equal(dst.astype(np.float64), src[k])
        except BaseException as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=caller, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not any(t.is_alive() for t in threads), "a caller is stuck"
    assert not errors, errors
