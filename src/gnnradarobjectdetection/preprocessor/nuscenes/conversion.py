"""Reference preprocessor/nuscenes/conversion.py:70-109 -> radargnn_b200.preprocessor.graph_construction."""
from radargnn_b200.preprocessor.graph_construction import build_geometric_graph, build_geometric_graphs  # noqa: F401

__all__ = ["build_geometric_graph", "build_geometric_graphs"]
