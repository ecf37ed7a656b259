"""Reference gnn/gnn_models.py:14-178 -> radargnn_b200.gnn.gnn_models."""
from radargnn_b200.gnn.gnn_models import DetNetBasic, get_mlp  # noqa: F401
from radargnn_b200.gnn._message_passing import BatchNorm, Linear  # noqa: F401

__all__ = ["DetNetBasic", "get_mlp"]
