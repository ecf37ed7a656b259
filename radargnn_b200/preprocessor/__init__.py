"""Mirror of the hot-path part of the reference's ``gnnradarobjectdetection.preprocessor``."""
from .configs import GraphConstructionConfiguration
from .radar_point_cloud import RadarPointCloud
from .graph_construction import (GraphConstructor, build_geometric_graph, build_geometric_graphs, collate_graph_data,
                                 create_graph_data)

__all__ = ["GraphConstructionConfiguration", "RadarPointCloud", "GraphConstructor", "build_geometric_graph",
           "build_geometric_graphs", "create_graph_data", "collate_graph_data"]
