"""Mirror of the hot-path part of the reference's ``gnnradarobjectdetection.gnn`` package."""
from .configs import GNNArchitectureConfig
from .mpnn_layers import MPNNConv, RadarPointGNNConv
from .gnn_models import DetNetBasic, get_mlp

__all__ = ["GNNArchitectureConfig", "MPNNConv", "RadarPointGNNConv", "DetNetBasic", "get_mlp"]
