# -*- coding: utf-8 -*-
"""
File I/O of the ``pyscenic aucell`` command (mirrors reference ``src/pyscenic/cli/utils.py``):

    load_exp_matrix   utils.py:114-157   csv / tsv (always); 10x MatrixMarket triplets (additive; always sparse);
                                         loom / h5ad when loompy / anndata are importable
    load_signatures   utils.py:205-227   gmt, dat (pickle), yaml (regulon records written by this package)
    save_matrix       utils.py:160-179   csv / tsv (loom when loompy is importable)

Unlike the reference, ``load_exp_matrix(..., return_sparse=True)`` is honoured for AUCell: a loom / h5ad matrix is
handed to the GPU as CSR without densifying.
"""
from __future__ import annotations

import pickle
from pathlib import PurePath
from typing import Sequence

import numpy as np
import pandas as pd

from ..genesig import GeneSignature, Regulon, openfile

__all__ = ["load_exp_matrix", "load_signatures", "save_matrix", "suffixes_to_separator", "guess_separator"]

ATTRIBUTE_NAME_CELL_IDENTIFIER = "CellID"
ATTRIBUTE_NAME_GENE = "Gene"

_MATRIX_SUFFIXES = (".csv", ".tsv", ".loom", ".h5ad", ".mtx")


def suffixes_to_separator(extension) -> str:
    if ".csv" in extension:
        return ","
    if ".tsv" in extension:
        return "\t"
    raise ValueError("cannot derive a column separator from suffixes {}".format(extension))


def _known_matrix(extension) -> bool:
    return any(s in extension for s in _MATRIX_SUFFIXES)


def _load_loom(fname, return_sparse, attr_cell, attr_gene):
    try:
        import loompy as lp
    except ImportError as e:  # pragma: no cover - optional dependency
        raise ValueError("reading loom files needs the loompy package: {}".format(e))
    with lp.connect(fname, mode="r", validate=False) as ds:
        genes = ds.ra[attr_gene]
        cells = ds.ca[attr_cell]
        if return_sparse:
            return ds.layers[""].sparse().T.tocsr(), genes, cells
        return pd.DataFrame(ds[:, :], index=genes, columns=cells).T


def _load_h5ad(fname, return_sparse):
    try:
        from anndata import read_h5ad
    except ImportError as e:  # pragma: no cover - optional dependency
        raise ValueError("reading h5ad files needs the anndata package: {}".format(e))
    import scipy.sparse as sp

    adata = read_h5ad(filename=fname, backed="r")
    x = adata.X[:] if hasattr(adata.X, "__getitem__") else adata.X
    if return_sparse:
        return sp.csr_matrix(x), adata.var_names.values, adata.obs_names.values
    dense = x.todense() if sp.issparse(x) else np.asarray(x)
    return pd.DataFrame(np.asarray(dense), index=adata.obs_names.values, columns=adata.var_names.values)


def _first_existing(directory, names):
    import os

    for n in names:
        p = os.path.join(directory, n)
        if os.path.exists(p):
            return p
    return None


def _load_10x_mtx(fname):
    """10x Genomics MatrixMarket triplet: ``matrix.mtx[.gz]`` (features x barcodes) with ``features.tsv[.gz]`` /
    ``genes.tsv`` and ``barcodes.tsv[.gz]`` next to it.  Returned as CSR cells x genes -- never densified, which is
    what makes 10^6-cell matrices loadable at all (the reference densifies every input, cli/utils.py:82-90,140-149)."""
    import os

    import scipy.io
    import scipy.sparse as sp

    d = os.path.dirname(os.path.abspath(fname))
    feat = _first_existing(d, ("features.tsv.gz", "features.tsv", "genes.tsv.gz", "genes.tsv"))
    barc = _first_existing(d, ("barcodes.tsv.gz", "barcodes.tsv"))
    if feat is None or barc is None:
        raise ValueError('"{}": features.tsv / genes.tsv and barcodes.tsv are expected next to the .mtx file.'.format(fname))
    ftab = pd.read_csv(feat, sep="\t", header=None)
    genes = ftab.iloc[:, 1 if ftab.shape[1] > 1 else 0].astype(str).values
    cells = pd.read_csv(barc, sep="\t", header=None).iloc[:, 0].astype(str).values
    m = sp.csr_matrix(scipy.io.mmread(fname).T)
    if m.shape != (len(cells), len(genes)):
        raise ValueError('"{}": matrix shape {} does not match {} barcodes x {} features.'.format(
            fname, m.shape[::-1], len(cells), len(genes)))
    return m, genes, cells


def load_exp_matrix(
    fname: str,
    transpose: bool = False,
    return_sparse: bool = False,
    attribute_name_cell_id: str = ATTRIBUTE_NAME_CELL_IDENTIFIER,
    attribute_name_gene: str = ATTRIBUTE_NAME_GENE,
):
    """
    Load an expression matrix (rows = cells x columns = genes).

    :return: a DataFrame, or ``(csr_matrix, gene_names, cell_names)`` when ``return_sparse`` and the file is loom/h5ad.
    """
    extension = PurePath(fname).suffixes
    if not _known_matrix(extension):
        raise ValueError('Unknown file format "{}".'.format(fname))
    if ".loom" in extension:
        return _load_loom(fname, return_sparse, attribute_name_cell_id, attribute_name_gene)
    if ".h5ad" in extension:
        return _load_h5ad(fname, return_sparse)
    if ".mtx" in extension:
        m, genes, cells = _load_10x_mtx(fname)
        if return_sparse:
            return m, genes, cells
        return pd.DataFrame(np.asarray(m.todense()), index=cells, columns=genes)
    df = pd.read_csv(fname, sep=suffixes_to_separator(extension), header=0, index_col=0)
    return df.T if transpose else df


def save_matrix(df: pd.DataFrame, fname: str, transpose: bool = False) -> None:
    """Save a matrix (rows = cells x columns = regulons) as csv / tsv (or a single-layer loom)."""
    extension = PurePath(fname).suffixes
    if not _known_matrix(extension):
        raise ValueError('Unknown file format "{}".'.format(fname))
    if ".loom" in extension:
        try:
            import loompy as lp
        except ImportError as e:  # pragma: no cover - optional dependency
            raise ValueError("writing loom files needs the loompy package: {}".format(e))
        lp.create(
            filename=fname,
            layers=df.T.values,
            row_attrs={ATTRIBUTE_NAME_GENE: df.columns.values.astype("str")},
            col_attrs={ATTRIBUTE_NAME_CELL_IDENTIFIER: df.index.values.astype("str")},
        )
        return
    if ".h5ad" in extension:
        raise ValueError("h5ad output needs an h5ad input; use csv/tsv/loom for a stand-alone AUC matrix")
    (df.T if transpose else df).to_csv(fname, sep=suffixes_to_separator(extension))


def guess_separator(fname: str) -> str:
    """Field separator of a GMT-like file: the first of tab, ';', ',' that yields at least three columns on every
    non-comment line."""
    with openfile(fname, "r") as f:
        lines = [ln.decode() if isinstance(ln, (bytes, bytearray)) else ln for ln in f.readlines()]
    body = [ln for ln in lines if ln.strip() and not ln.strip().startswith("#")]
    for sep in ("\t", ";", ","):
        if body and min(len(ln.split(sep)) for ln in body) >= 3:
            return sep
    raise ValueError('Unknown file format "{}".'.format(fname))


def save_to_yaml(signatures: Sequence[GeneSignature], fname: str) -> None:
    import yaml

    docs = []
    for s in signatures:
        d = {"name": s.name, "gene2weight": {g: float(w) for g, w in s.gene2weight.items()}}
        if isinstance(s, Regulon):
            d.update(transcription_factor=s.transcription_factor, context=sorted(s.context), score=float(s.score))
        docs.append(d)
    with openfile(fname, "w") as f:
        yaml.safe_dump(docs, f)


# ---- cisTarget motif-enrichment tables -> regulons ---------------------------------------------------------------
# `pyscenic aucell expr.loom reg.csv` is the standard invocation: the signatures file is the table `pyscenic ctx`
# writes.  Reference: cli/utils.py:205-227 (load_signatures), utils.py:438-456 (load_motifs), transform.py:382-536
# (df2regulons / _regulon4group).  Only what AUCell needs is kept: one regulon per (TF, activating / repressing), the
# union of the target genes of its enriched motifs with the largest weight per gene.
COLUMN_NAME_TF = "TF"
COLUMN_NAME_MOTIF_ID = "MotifID"
COLUMN_NAME_NES = "NES"
COLUMN_NAME_CONTEXT = "Context"
COLUMN_NAME_TARGET_GENES = "TargetGenes"
COLUMN_NAME_MOTIF_SIMILARITY_QVALUE = "MotifSimilarityQvalue"
COLUMN_NAME_ORTHOLOGOUS_IDENTITY = "OrthologousIdentity"
ACTIVATING_MODULE = "activating"
REPRESSING_MODULE = "repressing"


def _literal(text):
    """The table stores Python literals (``frozenset({...})`` contexts, ``[(gene, weight), ...]`` target genes); the
    reference ``eval``s them.  Here only literals are accepted."""
    import ast

    t = text.strip()
    if t.startswith("frozenset(") and t.endswith(")"):
        inner = t[len("frozenset("):-1].strip()
        return frozenset(ast.literal_eval(inner)) if inner else frozenset()
    return ast.literal_eval(t)


def load_motifs(fname: str, sep: str = ",") -> pd.DataFrame:
    """The enriched-motifs table of ``pyscenic ctx`` (two header rows, index (TF, MotifID))."""
    df = pd.read_csv(fname, sep=sep, index_col=[0, 1], header=[0, 1], skipinitialspace=True)
    df[("Enrichment", COLUMN_NAME_CONTEXT)] = df[("Enrichment", COLUMN_NAME_CONTEXT)].apply(_literal)
    df[("Enrichment", COLUMN_NAME_TARGET_GENES)] = df[("Enrichment", COLUMN_NAME_TARGET_GENES)].apply(_literal)
    return df


def df2regulons(df: pd.DataFrame) -> Sequence[Regulon]:
    """Regulons from a table of enriched motifs: rows are grouped per TF and interaction type; a group's target genes
    are united keeping the maximum weight per gene, its score is the best combined score of its motifs."""
    import math

    assert not df.empty, "Signatures dataframe is empty!"
    df = df.copy()
    if df.columns.nlevels == 2:
        df.columns = df.columns.droplevel(0)

    def score(nes, qval, ident):
        fraction = 1.0
        try:
            fraction = min(-math.log(qval, 10), 10) / 10 if not math.isnan(qval) else 1.0
        except ValueError:          # math domain error: q-value 0
            pass
        s = nes * fraction
        return s if math.isnan(ident) else s * ident

    regulons = []
    kinds = df[COLUMN_NAME_CONTEXT].apply(lambda ctx: REPRESSING_MODULE if REPRESSING_MODULE in ctx else ACTIVATING_MODULE)
    tfs = df.index.get_level_values(COLUMN_NAME_TF)
    for (tf_name, kind), grp in df.groupby([tfs, kinds.values], sort=True):
        context = frozenset([kind])
        name = "{}{}".format(tf_name, "(-)" if kind == REPRESSING_MODULE else "(+)")
        top = grp.sort_values(by=COLUMN_NAME_NES, ascending=False).head(1)
        motif_logo = "{}.png".format(top.index.get_level_values(COLUMN_NAME_MOTIF_ID)[0])
        merged = None
        for _, row in grp.iterrows():
            targets = row[COLUMN_NAME_TARGET_GENES]
            reg = Regulon(name=name, gene2weight=dict(targets) if not isinstance(targets, dict) else targets,
                          transcription_factor=tf_name, context=context,
                          score=score(float(row[COLUMN_NAME_NES]),
                                      float(row.get(COLUMN_NAME_MOTIF_SIMILARITY_QVALUE, float("nan"))),
                                      float(row.get(COLUMN_NAME_ORTHOLOGOUS_IDENTITY, float("nan")))))
            merged = reg if merged is None else merged.union(reg)
        regulons.append(merged.copy(context=frozenset(set(context) | {motif_logo})))
    return regulons


def _load_yaml(fname: str):
    """YAML signatures: this package's own plain layout (a list of mappings), or the reference's ``yaml.dump`` of
    signature objects (``!GeneSignature`` / ``!Regulon`` tagged mappings with ``genes`` / ``weights`` lists)."""
    import yaml

    def build(d):
        if "gene2weight" in d:
            g2w = d["gene2weight"]
        else:
            genes, weights = list(d.get("genes", ())), d.get("weights")
            g2w = dict(zip(genes, weights)) if weights is not None else genes
        if d.get("transcription_factor"):
            return Regulon(name=d["name"], gene2weight=g2w, transcription_factor=d["transcription_factor"],
                           context=frozenset(d.get("context", ()) or ()), score=d.get("score", 0.0) or 0.0)
        return GeneSignature(name=d["name"], gene2weight=g2w)

    class _Loader(yaml.SafeLoader):
        pass

    def tagged(loader, suffix, node):
        return build(loader.construct_mapping(node, deep=True))

    _Loader.add_multi_constructor("!", tagged)
    _Loader.add_multi_constructor("tag:yaml.org,2002:python/object", tagged)
    with openfile(fname, "r") as f:
        docs = yaml.load(f, Loader=_Loader)
    return [d if isinstance(d, GeneSignature) else build(d) for d in docs]


def load_signatures(fname: str) -> Sequence[GeneSignature]:
    """
    Load genes signatures from disk.

    Supported file formats are GMT, DAT (pickled), YAML or CSV / TSV (enriched motifs), as in the reference
    (``cli/utils.py:205-227``).

    :param fname: The name of the file that contains the signatures.
    :return: A list of gene signatures.
    """
    extension = PurePath(fname).suffixes
    if ".csv" in extension or ".tsv" in extension:
        return df2regulons(load_motifs(fname, sep=suffixes_to_separator(extension)))
    if ".yaml" in extension or ".yml" in extension:
        return _load_yaml(fname)
    if ".gmt" in extension:
        sep = guess_separator(fname)
        return GeneSignature.from_gmt(fname, field_separator=sep, gene_separator=sep)
    if ".dat" in extension:
        with openfile(fname, "rb") as f:
            return pickle.load(f)
    raise ValueError('Unknown file format "{}".'.format(fname))
