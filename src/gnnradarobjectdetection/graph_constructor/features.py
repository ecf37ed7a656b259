"""Reference graph_constructor/features.py:6-122 -> radargnn_b200.graph_constructor.features."""
from radargnn_b200.graph_constructor.features import get_En_equivariant_point_pair_metrics  # noqa: F401

__all__ = ["get_En_equivariant_point_pair_metrics"]
