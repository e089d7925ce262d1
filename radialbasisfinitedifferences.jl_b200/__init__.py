"""B200-native operator generation/application hot path of RadialBasisFiniteDifferences.jl.

The directory name is not an importable identifier; load it through the repo-root shim:  `import rbffd_b200`.
"""
from ._lib import ADVDIFF_COLLOCATED, AdvDiffParams, Options, RbffdError, build, exported_symbols, lib  # noqa: F401
from .api import (BoundaryConditions, Context, Operator, REFERENCE_OPS, bind_to_gpu_numa, calculateneighbors, default_context, generate_operator,  # noqa: F401
                  generate_operator_collocated, generate_raw, groups_from_index_sets, hyperviscosity_operator,
                  hyperviscosity_operator_collocated, make_options)
from . import lsq, mesh, nodes, sharding  # noqa: F401
from .sharding import PeerHalo, Shard, SlabShard, boundary_row_ranges, exchange_halo  # noqa: F401
