"""Mirror of the hot-path part of the reference's ``gnnradarobjectdetection.gnn`` package."""
from .configs import GNNArchitectureConfig
from .mpnn_layers import MPNNConv, RadarPointGNNConv
from .gnn_models import DetNetBasic, get_mlp
from ._autograd import detection_loss

__all__ = ["GNNArchitectureConfig", "MPNNConv", "RadarPointGNNConv", "DetNetBasic", "get_mlp", "detection_loss"]
