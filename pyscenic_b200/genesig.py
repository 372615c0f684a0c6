# -*- coding: utf-8 -*-
"""
Gene signatures and regulons: the input type of the AUCell boundary.

Mirrors the public surface of ``ctxcore.genesig.GeneSignature`` / ``Regulon`` that the AUCell path and its
callers touch (reference ``src/pyscenic/aucell.py:14,100,124,148,179``; built at
``tests/test_aucell.py:24-28,51``, ``src/pyscenic/transform.py:414-425``).  ``ctxcore`` is a third-party
dependency of the reference that is not available in this image, so this is an independent implementation of
the same record: an immutable ``name`` + ``gene2weight`` mapping with set algebra, GMT round-tripping and
``noweights()``.  The AUCell entry points are duck-typed (``name``, ``genes``, ``len()``, ``sig[gene]``,
``noweights()``), so real ``ctxcore`` objects are accepted as well.
"""
from __future__ import annotations

import gzip
import re
from collections.abc import Iterable, Mapping
from itertools import repeat
from operator import itemgetter
from types import MappingProxyType
from typing import FrozenSet, List, Tuple

__all__ = ["GeneSignature", "Regulon", "openfile"]


def openfile(filename: str, mode: str = "r"):
    """Open plain or gzip-compressed text files transparently."""
    if filename.endswith(".gz"):
        return gzip.open(filename, mode + "t" if "b" not in mode and "t" not in mode else mode)
    return open(filename, mode)


def _convert(genes) -> Mapping:
    """Accepted spellings of ``gene2weight``: mapping, iterable of ``(gene, weight)`` pairs, iterable of names
    (every weight 1.0)."""
    if isinstance(genes, Mapping):
        d = dict(genes)
    elif isinstance(genes, Iterable) and not isinstance(genes, (str, bytes)):
        items = list(genes)
        if all(isinstance(n, tuple) for n in items):
            d = dict(items)
        elif all(isinstance(n, str) for n in items):
            d = dict(zip(items, repeat(1.0)))
        else:
            raise ValueError("gene2weight must be a mapping, an iterable of (gene, weight) tuples or of gene names.")
    else:
        raise ValueError("gene2weight must be a mapping, an iterable of (gene, weight) tuples or of gene names.")
    return MappingProxyType(d)


def _rebuild(cls, fields):
    return cls(**fields)


class GeneSignature:
    """A named gene set with one weight per gene (immutable)."""

    __slots__ = ("_name", "_gene2weight", "_hash")

    def __init__(self, name: str, gene2weight):
        if not isinstance(name, str) or len(name) == 0:
            raise ValueError("A gene signature must have a non-empty name.")
        g2w = _convert(gene2weight)
        object.__setattr__(self, "_name", name)
        object.__setattr__(self, "_gene2weight", g2w)
        object.__setattr__(self, "_hash", None)

    def __setattr__(self, key, value):
        raise AttributeError("{} is immutable".format(type(self).__name__))

    # ------------------------------------------------------------------ record fields
    @property
    def name(self) -> str:
        return self._name

    @property
    def gene2weight(self) -> Mapping:
        return self._gene2weight

    @property
    def genes(self) -> Tuple[str, ...]:
        """Gene names, by descending weight (stable for equal weights)."""
        return tuple(map(itemgetter(0), sorted(self._gene2weight.items(), key=itemgetter(1), reverse=True)))

    @property
    def weights(self) -> Tuple[float, ...]:
        """Weights in the order of :attr:`genes`."""
        return tuple(map(itemgetter(1), sorted(self._gene2weight.items(), key=itemgetter(1), reverse=True)))

    @property
    def genes_weights(self) -> Tuple[Tuple[str, float], ...]:
        return tuple(zip(self.genes, self.weights))

    def __len__(self) -> int:
        return len(self._gene2weight)

    def __contains__(self, item) -> bool:
        return item in self._gene2weight

    def __getitem__(self, item) -> float:
        return self._gene2weight[item]

    def __iter__(self):
        return iter(self.genes)

    def __eq__(self, other) -> bool:
        return (
            type(other) is type(self)
            and self._fields() == other._fields()
        )

    def __hash__(self) -> int:
        if self._hash is None:
            object.__setattr__(self, "_hash", hash((self._name, frozenset(self._gene2weight.items()))))
        return self._hash

    def __repr__(self) -> str:
        return "{}(name={!r}, gene2weight=<{} genes>)".format(type(self).__name__, self._name, len(self))

    def _fields(self) -> dict:
        return {"name": self._name, "gene2weight": dict(self._gene2weight)}

    def __reduce__(self):
        # pickling (the reference's .dat signature files, cli/utils.py:223-225)
        return (_rebuild, (type(self), self._fields()))

    # ------------------------------------------------------------------ derivation
    def copy(self, **kwargs) -> "GeneSignature":
        fields = self._fields()
        fields.update(kwargs)
        return type(self)(**fields)

    def rename(self, name: str) -> "GeneSignature":
        return self.copy(name=name)

    def add(self, gene_symbol: str, weight: float = 1.0) -> "GeneSignature":
        d = dict(self._gene2weight)
        d[gene_symbol] = weight
        return self.copy(gene2weight=d)

    def noweights(self) -> "GeneSignature":
        """Same genes, every weight 1.0 (what ``aucell(..., noweights=True)`` applies, reference
        ``aucell.py:124,147-148``)."""
        return self.copy(gene2weight=self.genes)

    def head(self, n: int = 5) -> "GeneSignature":
        assert n >= 1, "n must be greater than or equal to one."
        return self.copy(gene2weight=self.genes_weights[:n])

    def _combined_name(self, other, sep: str) -> str:
        return self._name if self._name == other.name else "({} {} {})".format(self._name, sep, other.name)

    def union(self, other: "GeneSignature") -> "GeneSignature":
        """Union of genes; a gene present in both keeps the larger weight."""
        d = dict(self._gene2weight)
        for g, w in other.gene2weight.items():
            d[g] = max(w, d[g]) if g in d else w
        return self.copy(name=self._combined_name(other, "|"), gene2weight=d)

    def difference(self, other: "GeneSignature") -> "GeneSignature":
        d = {g: w for g, w in self._gene2weight.items() if g not in other.gene2weight}
        return self.copy(name=self._combined_name(other, "-"), gene2weight=d)

    def intersection(self, other: "GeneSignature") -> "GeneSignature":
        d = {g: max(w, other.gene2weight[g]) for g, w in self._gene2weight.items() if g in other.gene2weight}
        return self.copy(name=self._combined_name(other, "&"), gene2weight=d)

    def jaccard_index(self, other: "GeneSignature") -> float:
        ga, gb = set(self._gene2weight), set(other.gene2weight)
        return float(len(ga & gb)) / len(ga | gb)

    # ------------------------------------------------------------------ file formats
    @classmethod
    def from_gmt(cls, fname: str, field_separator: str = "\t", gene_separator: str = ",") -> List["GeneSignature"]:
        """GMT: column 0 name, column 1 description, remainder the genes (reference caller:
        ``src/pyscenic/cli/utils.py:220-222``, ``tests/test_aucell.py:22-28``)."""
        out = []
        with openfile(fname, "r") as fh:
            for line in fh:
                if isinstance(line, (bytes, bytearray)):
                    line = line.decode()
                if line.startswith("#") or not line.strip():
                    continue
                name, _, genes_str = re.split(field_separator, line.rstrip(), maxsplit=2)
                genes = genes_str.split(gene_separator)
                out.append(GeneSignature(name=name, gene2weight=genes))
        return out

    @classmethod
    def to_gmt(
        cls, fname: str, signatures: List["GeneSignature"], field_separator: str = "\t", gene_separator: str = ","
    ) -> None:
        with openfile(fname, "wt") as fh:
            for sig in signatures:
                genes = gene_separator.join(sig.genes)
                fh.write("{}{}{}{}{}\n".format(sig.name, field_separator, "", field_separator, genes))

    @classmethod
    def from_grp(cls, fname: str, name: str) -> "GeneSignature":
        with openfile(fname, "r") as fh:
            genes = [ln.strip() for ln in fh if ln.strip() and not ln.startswith("#")]
        return GeneSignature(name=name, gene2weight=genes)

    @classmethod
    def from_rnk(cls, fname: str, name: str, field_separator: str = ",") -> "GeneSignature":
        pairs = []
        with openfile(fname, "r") as fh:
            for ln in fh:
                if ln.startswith("#") or not ln.strip():
                    continue
                cols = re.split(field_separator, ln.rstrip())
                pairs.append((cols[0], float(cols[1])))
        return GeneSignature(name=name, gene2weight=pairs)


class Regulon(GeneSignature):
    """A gene signature that is the target set of one transcription factor.  Named ``TF(+)`` / ``TF(-)`` by the
    reference (``src/pyscenic/transform.py:409-415``); AUCell scores both kinds with identical arithmetic."""

    __slots__ = (
        "_gene2occurrence",
        "_transcription_factor",
        "_context",
        "_score",
        "_nes",
        "_orthologous_identity",
        "_similarity_qvalue",
        "_annotation",
    )

    def __init__(
        self,
        name: str,
        gene2weight,
        transcription_factor: str,
        gene2occurrence=None,
        context: FrozenSet[str] = frozenset(),
        score: float = 0.0,
        nes: float = 0.0,
        orthologous_identity: float = 0.0,
        similarity_qvalue: float = 0.0,
        annotation: str = "",
    ):
        super().__init__(name, gene2weight)
        if not isinstance(transcription_factor, str) or len(transcription_factor) == 0:
            raise ValueError("A regulon must have a transcription factor.")
        s = object.__setattr__
        s(self, "_gene2occurrence", _convert(gene2occurrence if gene2occurrence is not None else {}))
        s(self, "_transcription_factor", transcription_factor)
        s(self, "_context", frozenset(context))
        s(self, "_score", score)
        s(self, "_nes", nes)
        s(self, "_orthologous_identity", orthologous_identity)
        s(self, "_similarity_qvalue", similarity_qvalue)
        s(self, "_annotation", annotation)

    gene2occurrence = property(lambda self: self._gene2occurrence)
    transcription_factor = property(lambda self: self._transcription_factor)
    context = property(lambda self: self._context)
    score = property(lambda self: self._score)
    nes = property(lambda self: self._nes)
    orthologous_identity = property(lambda self: self._orthologous_identity)
    similarity_qvalue = property(lambda self: self._similarity_qvalue)
    annotation = property(lambda self: self._annotation)

    def _fields(self) -> dict:
        d = super()._fields()
        d.update(
            transcription_factor=self._transcription_factor,
            gene2occurrence=dict(self._gene2occurrence),
            context=self._context,
            score=self._score,
            nes=self._nes,
            orthologous_identity=self._orthologous_identity,
            similarity_qvalue=self._similarity_qvalue,
            annotation=self._annotation,
        )
        return d

    def __hash__(self) -> int:
        if self._hash is None:
            object.__setattr__(
                self, "_hash", hash((self._name, self._transcription_factor, frozenset(self._gene2weight.items())))
            )
        return self._hash

    def noweights(self) -> "Regulon":
        return self.copy(gene2weight=self.genes)

    def union(self, other: "GeneSignature") -> "Regulon":
        assert getattr(other, "transcription_factor", self._transcription_factor) == self._transcription_factor, (
            "Union of two regulons is only possible when same factor."
        )
        merged = super().union(other)
        occ = dict(self._gene2occurrence)
        for g, n in getattr(other, "gene2occurrence", {}).items():
            occ[g] = occ.get(g, 0) + n
        return merged.copy(
            gene2occurrence=occ,
            context=self._context.union(getattr(other, "context", frozenset())),
            score=max(self._score, getattr(other, "score", 0.0)),
            nes=max(self._nes, getattr(other, "nes", 0.0)),
        )
