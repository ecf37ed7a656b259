"""Reference preprocessor/radarscenes/dataset_creation.py:187-229 (GraphConstructor.build_geometric_graph) and
:786-814 (create_graph_data) -> radargnn_b200.preprocessor.graph_construction.  The dataset readers of that
module stay with the reference (out of scope)."""
from radargnn_b200.preprocessor.graph_construction import GraphConstructor, create_graph_data  # noqa: F401

__all__ = ["GraphConstructor", "create_graph_data"]
