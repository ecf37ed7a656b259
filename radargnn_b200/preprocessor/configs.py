"""Mirror of reference preprocessor/configs.py:4-26 (argument object of build_geometric_graph)."""
from dataclasses import dataclass


@dataclass
class GraphConstructionConfiguration:
    """All settings required for creating a graph from a point cloud."""

    graph_construction_algorithm: str
    graph_construction_settings: dict

    node_features: list
    edge_features: list
    edge_mode: str

    distance_definition: str

    def __post_init__(self):
        if self.graph_construction_algorithm == "knn":
            self.k = self.graph_construction_settings.get("k")
            self.r = None
        elif self.graph_construction_algorithm == "radius":
            self.r = self.graph_construction_settings.get("r")
            self.k = None
        else:
            raise Exception("Invalid graph construction algorithm selected")
