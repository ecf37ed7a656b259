"""Reference gnn/configs.py:4-30 -> radargnn_b200.gnn.configs."""
from radargnn_b200.gnn.configs import GNNArchitectureConfig  # noqa: F401

__all__ = ["GNNArchitectureConfig"]
