"""vlfs_b200 - B200-native assembly + solve path behind the Gridap surface used by
MonolithicFEMVLFS.jl (`AffineFEOperator` / `SparseMatrixAssembler` / `solve` / `Newmark`).

The package directory is named `monolithicfemvlfs.jl_b200`; import it as `vlfs_b200`
(see `vlfs_b200.py` at the repository root).  Host code marshals tables; all arithmetic is
in `libvlfs_b200.so` (hand-written sm_100a CUDA).  No CPU fallback exists.
"""
from . import _lib
from ._lib import VlfsError, LIB_PATH, EXPORTS
from .geometry import (CartesianDiscreteModel, DiscreteModelFromFile, Interior, Boundary,
                       Triangulation, Skeleton, add_tag_from_tags)
from .fe import (ReferenceFE, TestFESpace, TrialFESpace, FESpace, MultiFieldFESpace, Measure,
                 lagrangian)
from .assembler import B200System

__all__ = ["CartesianDiscreteModel", "DiscreteModelFromFile", "Interior", "Boundary",
           "Triangulation", "Skeleton", "add_tag_from_tags", "ReferenceFE", "TestFESpace",
           "TrialFESpace", "FESpace", "MultiFieldFESpace", "Measure", "lagrangian",
           "B200System", "VlfsError", "LIB_PATH", "EXPORTS"]
