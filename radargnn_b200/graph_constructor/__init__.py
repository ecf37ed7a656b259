"""Mirror of the reference's ``gnnradarobjectdetection.graph_constructor`` package."""
from .graph import Graph, GeometricGraph
from .features import get_En_equivariant_point_pair_metrics

__all__ = ["Graph", "GeometricGraph", "get_En_equivariant_point_pair_metrics"]
