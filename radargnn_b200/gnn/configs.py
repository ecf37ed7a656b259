"""Mirror of reference gnn/configs.py:4-30 (the model's parameter object; field names kept)."""
from dataclasses import dataclass


@dataclass
class GNNArchitectureConfig:
    """Possible GNN model architecture configurations."""

    # initial node and edge feature dimension
    node_feature_dimension: int
    edge_feature_dimension: int

    # layers for graph convolution and detection head
    conv_layer_dimensions: list
    classification_head_layer_dimensions: list
    regression_head_layer_dimensions: list

    # layers for initial node and edge feature embedding MLPs
    initial_node_feature_embedding: bool = False
    initial_edge_feature_embedding: bool = False
    node_feature_embedding_layer_dimensions: list = None
    edge_feature_embedding_layer_dimensions: list = None
    conv_layer_type: str = "MPNNConv"

    # configuration for graph convolution layers
    batch_norm_in_mlps: bool = True
    conv_pre_mlp_layer_number: int = 1
    conv_post_mlp_layer_number: int = 1
    conv_use_edge_encoder: bool = False
    aggregation_function: str = "max"
