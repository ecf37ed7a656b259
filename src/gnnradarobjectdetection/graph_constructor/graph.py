"""Reference graph_constructor/graph.py:10-302 -> radargnn_b200.graph_constructor.graph."""
from radargnn_b200.graph_constructor.graph import Graph, GeometricGraph  # noqa: F401

__all__ = ["Graph", "GeometricGraph"]
