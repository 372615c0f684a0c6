# -*- coding: utf-8 -*-
"""
pyscenic_b200 -- B200-native (sm_100a) AUCell: a drop-in for the ``pyscenic.aucell`` step of aertslab/pySCENIC.

    from pyscenic_b200.aucell import aucell, aucell4r, create_rankings, derive_auc_threshold
    from pyscenic_b200.genesig import GeneSignature, Regulon

The CUDA kernels live in ``pyscenic_b200/csrc`` behind the C-ABI declared in ``include/aucell_b200.h``; the
Python modules mirror the reference's own interface for this path.  No CPU fallback exists.
"""
import logging

from .genesig import GeneSignature, Regulon  # noqa: F401

__version__ = "0.1.0"

LOGGER = logging.getLogger(__name__)
LOGGER.addHandler(logging.NullHandler())
