# -*- coding: utf-8 -*-
"""`pyscenic aucell` command: argument surface and file I/O (CPU), end-to-end run against the oracle (GPU)."""
import os
import pickle

import numpy as np
import pandas as pd
import pytest

from oracle import aucell_oracle as O
from pyscenic_b200.cli import pyscenic as cli
from pyscenic_b200.cli.utils import (guess_separator, load_exp_matrix, load_signatures, save_matrix, save_to_yaml)
from pyscenic_b200.genesig import GeneSignature, Regulon


def _write_inputs(tmp_path, n_cells=40, n_genes=300, seed=0):
    rs = np.random.RandomState(seed)
    X = rs.poisson(0.6, size=(n_cells, n_genes)).astype(np.float64)
    df = pd.DataFrame(X, index=["cell%d" % i for i in range(n_cells)], columns=["g%d" % j for j in range(n_genes)])
    expr = str(tmp_path / "expr.csv")
    df.to_csv(expr)
    sigs = [GeneSignature("sig%d" % k, ["g%d" % g for g in rs.choice(n_genes, 30, replace=False)]) for k in range(6)]
    gmt = str(tmp_path / "sigs.gmt")
    GeneSignature.to_gmt(gmt, sigs, field_separator="\t", gene_separator="\t")
    return df, sigs, expr, gmt


def test_parser_has_the_reference_flag_surface(tmp_path):
    df, sigs, expr, gmt = _write_inputs(tmp_path)
    p = cli.create_argument_parser()
    a = p.parse_args(["aucell", expr, gmt])
    # defaults of the reference (cli/pyscenic.py:340-360, 650-708): unweighted, thr 0.05, seed None, stdout
    assert a.auc_threshold == 0.05 and a.weights is None and a.seed is None and a.transpose is None
    import sys

    assert a.output is sys.stdout and a.rank_threshold == 5000 and a.nes_threshold == 3.0
    assert a.cell_id_attribute == "CellID" and a.gene_attribute == "Gene" and a.sparse is False
    assert a.func is cli.aucell_command
    a = p.parse_args(["aucell", expr, gmt, "-o", str(tmp_path / "o.csv"), "-t", "-w", "--num_workers", "3", "--seed",
                      "7", "--auc_threshold", "0.1", "--sparse"])
    assert a.transpose == "yes" and a.weights == "yes" and a.num_workers == 3 and a.seed == 7 and a.sparse is True
    # @args file support
    argsfile = tmp_path / "args.txt"
    argsfile.write_text("aucell\n%s\n%s\n--seed\n5\n" % (expr, gmt))
    assert p.parse_args(["@" + str(argsfile)]).seed == 5


def _write_ctx_table(path, sep=","):
    """A motif-enrichment table laid out like `pyscenic ctx` writes it (two header rows, (TF, MotifID) index, Python
    literals in the Context / TargetGenes cells): the usual signatures file of `pyscenic aucell`."""
    nan = float("nan")
    rows = [("TFa", "m1", 0.1, 3.5, 1e-4, nan, "ann", frozenset({"activating", "weight>75.0%"}), [("g1", 2.0), ("g2", 1.0)], 10),
            ("TFa", "m2", 0.2, 4.5, nan, 0.8, "ann", frozenset({"activating", "top50"}), [("g2", 3.0), ("g3", 0.5)], 12),
            ("TFa", "m3", 0.2, 3.1, nan, nan, "ann", frozenset({"repressing", "top50"}), [("g9", 1.5), ("g7", 0.25)], 12),
            ("TFb", "m4", 0.3, 5.0, 0.0, nan, "ann", frozenset({"activating"}), [("g4", 1.0), ("g1", 4.0)], 3)]
    idx = pd.MultiIndex.from_tuples([(r[0], r[1]) for r in rows], names=["TF", "MotifID"])
    cols = pd.MultiIndex.from_tuples([("Enrichment", c) for c in ["AUC", "NES", "MotifSimilarityQvalue", "OrthologousIdentity",
                                                                  "Annotation", "Context", "TargetGenes", "RankAtMax"]])
    pd.DataFrame([r[2:] for r in rows], index=idx, columns=cols).to_csv(path, sep=sep)


def test_signatures_from_a_ctx_motif_table_and_reference_yaml(tmp_path):
    """reference cli/utils.py:205-227: csv / tsv signatures are cisTarget tables (-> df2regulons); yaml files may hold the
    reference's tagged signature objects."""
    for ext, sep in ((".csv", ","), (".tsv", "\t")):
        f = str(tmp_path / ("reg" + ext))
        _write_ctx_table(f, sep)
        regs = load_signatures(f)
        assert [r.name for r in regs] == ["TFa(+)", "TFa(-)", "TFb(+)"]
        assert all(isinstance(r, Regulon) for r in regs)
        # union of the motifs' target genes, the largest weight per gene
        assert dict(regs[0].gene2weight) == {"g1": 2.0, "g2": 3.0, "g3": 0.5} and regs[0].transcription_factor == "TFa"
        assert dict(regs[1].gene2weight) == {"g9": 1.5, "g7": 0.25} and "repressing" in regs[1].context
        assert "m2.png" in regs[0].context and abs(regs[0].score - 4.5 * 0.8) < 1e-12     # best combined score
        assert abs(regs[2].score - 5.0) < 1e-12                                           # q-value 0: no penalty
    yml = tmp_path / "ref.yaml"
    yml.write_text("- !GeneSignature\n  name: s1\n  genes: [g1, g2]\n  weights: [1.0, 2.5]\n"
                   "- !Regulon\n  name: TFz(+)\n  genes: [g3]\n  weights: [0.75]\n  transcription_factor: TFz\n"
                   "  context: [activating]\n  score: 2.0\n")
    back = load_signatures(str(yml))
    assert back[0].name == "s1" and back[0]["g2"] == 2.5 and not isinstance(back[0], Regulon)
    assert isinstance(back[1], Regulon) and back[1].transcription_factor == "TFz" and back[1]["g3"] == 0.75


def test_matrix_and_signature_io(tmp_path):
    df, sigs, expr, gmt = _write_inputs(tmp_path)
    back = load_exp_matrix(expr)
    assert back.shape == df.shape and list(back.columns) == list(df.columns) and np.array_equal(back.values, df.values)
    df.T.to_csv(str(tmp_path / "t.tsv"), sep="\t")
    assert np.array_equal(load_exp_matrix(str(tmp_path / "t.tsv"), transpose=True).values, df.values)
    with pytest.raises(ValueError):
        load_exp_matrix(str(tmp_path / "expr.parquet"))
    assert guess_separator(gmt) == "\t"
    got = load_signatures(gmt)
    assert [s.name for s in got] == [s.name for s in sigs] and set(got[0].genes) == set(sigs[0].genes)
    dat = str(tmp_path / "sigs.dat")
    pickle.dump(sigs, open(dat, "wb"))
    assert load_signatures(dat) == sigs
    regs = [Regulon("TF1(+)", {"g1": 0.5, "g2": 2.0}, transcription_factor="TF1", context=frozenset(["activating"]))]
    yml = str(tmp_path / "regs.yaml")
    save_to_yaml(regs + sigs[:1], yml)
    back = load_signatures(yml)
    assert isinstance(back[0], Regulon) and back[0]["g2"] == 2.0 and back[0].transcription_factor == "TF1"
    assert back[1] == sigs[0]
    with pytest.raises(ValueError):
        load_signatures(str(tmp_path / "x.json"))
    out = str(tmp_path / "m.tsv")
    save_matrix(df.iloc[:3, :4], out, transpose=True)
    assert pd.read_csv(out, sep="\t", index_col=0).shape == (4, 3)
    with pytest.raises(ValueError):
        save_matrix(df, str(tmp_path / "m.xlsx"))


def test_unknown_format_exits_with_status_1(tmp_path):
    df, sigs, expr, gmt = _write_inputs(tmp_path)
    bad = tmp_path / "expr.parquet"
    bad.write_text("x")
    with pytest.raises(SystemExit) as e:
        cli.main(["aucell", str(bad), gmt])
    assert e.value.code == 1


def test_10x_mtx_triplet_loads_sparse(tmp_path):
    import gzip

    import scipy.io
    import scipy.sparse as sp

    df, sigs, expr, gmt = _write_inputs(tmp_path, n_cells=25, n_genes=60)
    d = tmp_path / "filtered_feature_bc_matrix"
    d.mkdir()
    scipy.io.mmwrite(str(d / "matrix.mtx"), sp.coo_matrix(df.values.T))          # features x barcodes, like 10x
    with gzip.open(str(d / "features.tsv.gz"), "wt") as f:
        f.write("".join("ENSG%05d\t%s\tGene Expression\n" % (j, g) for j, g in enumerate(df.columns)))
    (d / "barcodes.tsv").write_text("".join("%s\n" % c for c in df.index))
    m, genes, cells = load_exp_matrix(str(d / "matrix.mtx"), return_sparse=True)
    assert sp.isspmatrix_csr(m) and m.shape == df.shape
    assert list(genes) == list(df.columns) and list(cells) == list(df.index)
    assert np.array_equal(np.asarray(m.todense()), df.values)
    dense = load_exp_matrix(str(d / "matrix.mtx"))
    assert np.array_equal(dense.values, df.values) and list(dense.columns) == list(df.columns)
    (d / "barcodes.tsv").unlink()
    with pytest.raises(ValueError):
        load_exp_matrix(str(d / "matrix.mtx"))


@pytest.mark.gpu
def test_cli_end_to_end_matches_oracle(tmp_path):
    df, sigs, expr, gmt = _write_inputs(tmp_path, n_cells=64, n_genes=500)
    out = str(tmp_path / "auc.csv")
    cli.main(["aucell", expr, gmt, "-o", out, "--seed", "42", "--auc_threshold", "0.1"])
    got = pd.read_csv(out, index_col=0)
    want = O.aucell(df, sigs, auc_threshold=0.1, noweights=True, seed=42, num_workers=2)
    assert list(got.index) == list(want.index) and list(got.columns) == list(want.columns)
    assert np.allclose(got.values, want.values, rtol=1e-12, atol=0)     # csv text round trip of exact values
    # transposed input and output, weighted flag, two shards on the same GPU
    df.T.to_csv(str(tmp_path / "t.csv"))
    out2 = str(tmp_path / "auc_t.tsv")
    cli.main(["aucell", str(tmp_path / "t.csv"), gmt, "-t", "-w", "-o", out2, "--seed", "42", "--auc_threshold", "0.1",
              "--devices", "0,0"])
    got2 = pd.read_csv(out2, sep="\t", index_col=0)
    assert got2.shape == (len(sigs), df.shape[0])
    assert np.allclose(got2.T.values, want.values, rtol=1e-12, atol=0)
    # 10x MatrixMarket input goes to the GPU as CSR
    import scipy.io
    import scipy.sparse as sp

    d = tmp_path / "tenx"
    d.mkdir()
    scipy.io.mmwrite(str(d / "matrix.mtx"), sp.coo_matrix(df.values.T))
    (d / "genes.tsv").write_text("".join("id%d\t%s\n" % (j, g) for j, g in enumerate(df.columns)))
    (d / "barcodes.tsv").write_text("".join("%s\n" % c for c in df.index))
    out3 = str(tmp_path / "auc_10x.csv")
    cli.main(["aucell", str(d / "matrix.mtx"), gmt, "-o", out3, "--seed", "42", "--auc_threshold", "0.1"])
    got3 = pd.read_csv(out3, index_col=0)
    assert list(got3.index) == list(want.index)
    assert np.allclose(got3.values, want.values, rtol=1e-12, atol=0)
