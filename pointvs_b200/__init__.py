"""B200-native EGNN scoring path for PointVS."""
from .egnn import (EGNNLayer, PygLinearPass, SartorrasEGNN,  # noqa: F401
                   MultitaskSatorrasEGNN, PNNGeometricBase)
from .graph import (CSRGraph, PackedBatch, generate_edges,  # noqa: F401
                    radius_graph_batch, csr_from_edge_index)
from .base import PointNeuralNetworkBase, DEVICE  # noqa: F401

__all__ = ['EGNNLayer', 'PygLinearPass', 'SartorrasEGNN',
           'MultitaskSatorrasEGNN', 'PNNGeometricBase', 'CSRGraph',
           'PackedBatch', 'generate_edges', 'radius_graph_batch',
           'csr_from_edge_index', 'PointNeuralNetworkBase', 'DEVICE']
