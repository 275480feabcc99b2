"""sirgcn_b200 — B200-native SIR-GCN convolution (drop-in for the reference's models/conv.py).

The directory is named ``sir-gcn_b200``; import it as ``sirgcn_b200`` through the shim module
``sirgcn_b200.py`` at the repository root (or via ``models.conv`` for the reference's scripts).
"""
from . import _lib
from .conv import SIRConv, SIREConv, SIRConvBase, SIREConvBase, classify_activation
from .function import EdgeAggregate, GatherAdd, SegmentReduce
from .graph import CompressedRows, DropEdge, Graph, as_graph

__all__ = ["SIRConv", "SIREConv", "SIRConvBase", "SIREConvBase", "Graph", "CompressedRows", "DropEdge", "as_graph",
           "EdgeAggregate", "GatherAdd", "SegmentReduce", "classify_activation"]
