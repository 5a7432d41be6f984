"""B200-native SageBlock hot path of aledima00/sldm-gnn.

    from sldm_gnn_b200 import SageBlock      # drop-in for src/models/blocks/sageblock.py

Importing this package loads libsldm_sage.so (hand-written sm_100a CUDA behind a
C-ABI, include/sldm_sage.h) and fails if it has not been built.  There is no CPU
fallback.
"""
from . import _lib  # noqa: F401  (raises ImportError if the CUDA library is missing)
from . import ops
from .ops import Csr, build_csr, segment_reduce
from .sageblock import SageBlock, SageConvParams, GraphedSageBlock
from .shim import install_reference_shim
from .readout import global_mean_pool, global_max_pool, global_mean_max_pool
from .collate import GraphData, GraphBatch, collate
from .map_attention import MapSpatialAttention
from .edges import build_proximity_edges
from .grusage import GruSage, MapEncoder, MapZscoreNorm, GraphedGruSage, GraphedTrainStep

__all__ = ["SageBlock", "SageConvParams", "GraphedSageBlock", "Csr", "build_csr", "segment_reduce", "ops", "install_reference_shim",
           "global_mean_pool", "global_max_pool", "global_mean_max_pool", "GraphData", "GraphBatch", "collate",
           "MapSpatialAttention", "build_proximity_edges", "GruSage", "MapEncoder", "MapZscoreNorm", "GraphedGruSage", "GraphedTrainStep"]
