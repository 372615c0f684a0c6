# -*- coding: utf-8 -*-
"""
``pyscenic aucell`` on B200 GPUs -- same positional arguments, flags and output behaviour as the reference command
(reference ``src/pyscenic/cli/pyscenic.py``: command ``:265-337``, parser ``:650-708``, shared argument groups
``:340-360`` and ``:454-475``).

    python -m pyscenic_b200.cli.pyscenic aucell expr.csv sigs.gmt -o auc.csv --seed 42

Differences, all additive: ``--sparse`` is honoured for loom / h5ad inputs (CSR goes to the GPU without densifying);
``--devices 0,1,...`` shards the cells over several GPUs.  ``--num_workers``, ``--rank_threshold`` and
``--nes_threshold`` are accepted and, as in the reference, do not influence the AUCell result
(``--num_workers 1`` selects the reference's label-sorted output of its serial branch).
"""
from __future__ import annotations

import argparse
import logging
import sys
from multiprocessing import cpu_count
from pathlib import PurePath
from shutil import copyfile

from ..aucell import aucell
from .utils import (ATTRIBUTE_NAME_CELL_IDENTIFIER, ATTRIBUTE_NAME_GENE, load_exp_matrix, load_signatures,
                    save_matrix)

LOGGER = logging.getLogger("pyscenic_b200.cli")


def _setup_logging():
    if not LOGGER.handlers:
        h = logging.StreamHandler(stream=sys.stderr)
        h.setFormatter(logging.Formatter("\n%(asctime)s - %(name)s - %(levelname)s - %(message)s"))
        root = logging.getLogger("pyscenic_b200")
        root.addHandler(h)
        root.setLevel(logging.INFO)


def aucell_command(args) -> None:
    """Calculate regulon enrichment (as AUC values) for cells."""
    _setup_logging()
    LOGGER.info("Loading expression matrix.")
    suffixes = PurePath(args.expression_mtx_fname.name).suffixes
    # 10x MatrixMarket triplets are always consumed sparse; loom / h5ad when --sparse is given
    sparse_in = ".mtx" in suffixes or (bool(args.sparse) and any(s in suffixes for s in (".loom", ".h5ad")))
    try:
        ex = load_exp_matrix(args.expression_mtx_fname.name, args.transpose == "yes", sparse_in,
                             args.cell_id_attribute, args.gene_attribute)
    except ValueError as e:
        LOGGER.error(e)
        sys.exit(1)

    LOGGER.info("Loading gene signatures.")
    try:
        signatures = load_signatures(args.signatures_fname.name)
    except ValueError as e:
        LOGGER.error(e)
        sys.exit(1)

    LOGGER.info("Calculating cellular enrichment.")
    devices = None if not args.devices else [int(d) for d in str(args.devices).split(",")]
    common = dict(auc_threshold=args.auc_threshold, noweights=(args.weights != "yes"), seed=args.seed,
                  num_workers=args.num_workers, devices=devices)
    if sparse_in:
        mtx, genes, cells = ex
        auc_mtx = aucell(mtx, signatures, genes=genes, cells=cells, **common)
    else:
        auc_mtx = aucell(ex, signatures, **common)

    LOGGER.info("Writing results to file.")
    extension = PurePath(args.output.name).suffixes
    transpose = args.transpose == "yes"
    if ".loom" in extension or ".h5ad" in extension:
        # the reference copies the input container and appends the AUC matrix (plus binarised regulons for loom);
        # that post-processing (binarization, embeddings) is downstream of this package: write the matrix itself
        try:
            if ".loom" in extension:
                save_matrix(auc_mtx, args.output.name, transpose)
            else:
                from anndata import read_h5ad  # noqa: F401 - optional dependency

                if ".h5ad" not in PurePath(args.expression_mtx_fname.name).suffixes:
                    raise ValueError("Expression matrix should be provided in the h5ad (anndata) file format.")
                copyfile(args.expression_mtx_fname.name, args.output.name)
                adata = read_h5ad(args.output.name)
                adata.obsm["X_aucell"] = auc_mtx.loc[adata.obs_names].values
                adata.uns["aucell_regulons"] = list(auc_mtx.columns)
                adata.write(args.output.name)
        except (ValueError, ImportError, OSError) as e:
            LOGGER.error(e)
            sys.exit(1)
    elif args.output.name == "<stdout>":
        (auc_mtx.T if transpose else auc_mtx).to_csv(args.output)
    else:
        try:
            save_matrix(auc_mtx, args.output.name, transpose)
        except ValueError as e:
            LOGGER.error(e)
            sys.exit(1)


def add_recovery_parameters(parser):
    group = parser.add_argument_group("motif enrichment arguments")
    group.add_argument("--rank_threshold", type=int, default=5000,
                       help="The rank threshold used for deriving the target genes of an enriched motif (default: 5000).")
    group.add_argument("--auc_threshold", type=float, default=0.05,
                       help="The threshold used for calculating the AUC of a feature as fraction of ranked genes "
                            "(default: 0.05).")
    group.add_argument("--nes_threshold", type=float, default=3.0,
                       help="The Normalized Enrichment Score (NES) threshold for finding enriched features (default: 3.0).")
    return parser


def add_loom_parameters(parser):
    group = parser.add_argument_group("loom file arguments")
    group.add_argument("--cell_id_attribute", type=str, default=ATTRIBUTE_NAME_CELL_IDENTIFIER,
                       help="The name of the column attribute that specifies the identifiers of the cells in the loom file.")
    group.add_argument("--gene_attribute", type=str, default=ATTRIBUTE_NAME_GENE,
                       help="The name of the row attribute that specifies the gene symbols in the loom file.")
    group.add_argument("--sparse", action="store_const", const=True, default=False,
                       help="If set, load the expression data as a sparse matrix (loom / h5ad inputs).")
    return parser


def create_argument_parser():
    parser = argparse.ArgumentParser(
        prog="pyscenic",
        description="Single-Cell rEgulatory Network Inference and Clustering -- AUCell step on NVIDIA B200",
        fromfile_prefix_chars="@",
        add_help=True,
        epilog="Arguments can be read from file using a @args.txt construct.",
    )
    subparsers = parser.add_subparsers(help="sub-command help")
    p = subparsers.add_parser("aucell", help="Quantify activity of gene signatures across single cells.")
    p.add_argument("expression_mtx_fname", type=argparse.FileType("r"),
                   help="The name of the file that contains the expression matrix for the single cell experiment."
                        " csv/tsv (rows=cells x columns=genes), loom (rows=genes x columns=cells) or h5ad.")
    p.add_argument("signatures_fname", type=argparse.FileType("r"),
                   help="The name of the file that contains the gene signatures (gmt, yaml or dat).")
    p.add_argument("-o", "--output", type=argparse.FileType("w"), default=sys.stdout,
                   help="Output file/stream, a matrix of AUC values (csv, tsv or loom).")
    p.add_argument("-t", "--transpose", action="store_const", const="yes",
                   help="Transpose the expression matrix if supplied as csv (rows=genes x columns=cells).")
    p.add_argument("-w", "--weights", action="store_const", const="yes",
                   help="Use weights associated with genes in recovery analysis.")
    p.add_argument("--num_workers", type=int, default=cpu_count(),
                   help="Accepted for compatibility with the reference (default: {}).".format(cpu_count()))
    p.add_argument("--seed", type=int, required=False, default=None,
                   help="Seed for the expression matrix ranking step. The default is to use a random seed.")
    p.add_argument("--devices", type=str, default=None, help="Comma-separated GPU indices to shard the cells over.")
    add_recovery_parameters(p)
    add_loom_parameters(p)
    p.set_defaults(func=aucell_command)
    return parser


def main(argv=None):
    parser = create_argument_parser()
    args = parser.parse_args(args=argv)
    if not hasattr(args, "func"):
        parser.print_help()
    else:
        args.func(args)


if __name__ == "__main__":
    main()
