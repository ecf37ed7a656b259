"""Reference gnn/mpnn_layers.py:11-184 -> radargnn_b200.gnn.mpnn_layers (same class names and arguments)."""
from radargnn_b200.gnn.mpnn_layers import MPNNConv, RadarPointGNNConv  # noqa: F401
from radargnn_b200.gnn._message_passing import Linear, MessagePassing, reset  # noqa: F401

__all__ = ["MPNNConv", "RadarPointGNNConv"]
