"""Mirror of reference gnn/gnn_models.py: ``DetNetBasic`` and ``get_mlp`` with the same module /
parameter names (``convs.{i}.pre_mlp.{j}.weight``, ``batch_norms.{i}.module.*``, ...), so the
published state_dicts load; all arithmetic runs in the CUDA kernels."""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch
from torch.nn import ModuleList, ReLU, Sequential

from .. import ops
from ._message_passing import BatchNorm, Linear
from .configs import GNNArchitectureConfig
from .mpnn_layers import MPNNConv, RadarPointGNNConv


def _run_mlp(mlp: Sequential, x: torch.Tensor) -> torch.Tensor:
    """``get_mlp`` Sequential on the CUDA kernels: a ReLU is fused into the Linear (or folded
    into the BatchNorm kernel) next to it."""
    mods = list(mlp)
    relu_pending = False
    i = 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, ReLU):
            relu_pending = True
        elif isinstance(m, BatchNorm):
            if relu_pending:
                x = torch.relu(x)
                relu_pending = False
            fuse = i + 1 < len(mods) and isinstance(mods[i + 1], ReLU)
            x = m(x, relu=fuse)
            if fuse:
                i += 1
        else:
            x = m(x, relu_input=relu_pending)
            relu_pending = False
        i += 1
    if relu_pending:
        x = torch.relu(x)
    return x


class DetNetBasic(torch.nn.Module):
    """GNN for end to end object detection and semantic segmentation (reference
    gnn_models.py:14-134): optional node / edge embedding MLPs, a stack of graph convolutions each
    followed by BatchNorm + ReLU, and the classification / box-regression heads."""

    def __init__(self, config: GNNArchitectureConfig):
        super().__init__()

        self.batch_norm_mlps = config.batch_norm_in_mlps

        self.node_feat_dim = config.node_feature_dimension
        self.edge_feat_dim = config.edge_feature_dimension

        self.conv_layer_dimensions = config.conv_layer_dimensions
        self.initial_node_feature_embedding = config.initial_node_feature_embedding
        self.initial_edge_feature_embedding = config.initial_edge_feature_embedding

        self.conv_pre_mlp_layers = config.conv_pre_mlp_layer_number
        self.conv_post_mlp_layers = config.conv_post_mlp_layer_number
        self.conv_use_edge_encoder = config.conv_use_edge_encoder
        self.aggregation = config.aggregation_function
        self.conv_layer_type = config.conv_layer_type

        if config.initial_node_feature_embedding:
            layer_dimensions = config.node_feature_embedding_layer_dimensions[:-1]
            out_dim = config.node_feature_embedding_layer_dimensions[-1]
            self.node_emb_mlp = get_mlp(self.node_feat_dim, out_dim, layer_dimensions, self.batch_norm_mlps)
            self.node_feat_dim = out_dim

        if config.initial_edge_feature_embedding:
            layer_dimensions = config.edge_feature_embedding_layer_dimensions[:-1]
            out_dim = config.edge_feature_embedding_layer_dimensions[-1]
            self.edge_emb_mlp = get_mlp(self.edge_feat_dim, out_dim, layer_dimensions, self.batch_norm_mlps)
            self.edge_feat_dim = out_dim

        self.convs = ModuleList()
        self.batch_norms = ModuleList()

        layer_dim = self.conv_layer_dimensions[0]
        if config.conv_layer_type == "MPNNConv":
            conv = MPNNConv(self.node_feat_dim, layer_dim, self.edge_feat_dim, aggr=self.aggregation,
                            pre_layers=self.conv_pre_mlp_layers, post_layers=self.conv_post_mlp_layers,
                            use_edge_encoder=self.conv_use_edge_encoder)
        elif config.conv_layer_type == "RadarPointGNNConv":
            conv = RadarPointGNNConv(self.node_feat_dim, self.edge_feat_dim, aggr=self.aggregation,
                                     pre_layers=self.conv_pre_mlp_layers, post_layers=self.conv_post_mlp_layers)
            layer_dim = self.node_feat_dim
        else:
            raise Exception(
                f"{config.conv_layer_type} is invalid GNN conv layer type. Chose either MPNNConv or RadarPointGNNConv")

        self.convs.append(conv)
        self.batch_norms.append(BatchNorm(layer_dim))

        for next_layer_dim in self.conv_layer_dimensions[1:]:
            if config.conv_layer_type == "MPNNConv":
                conv = MPNNConv(layer_dim, next_layer_dim, self.edge_feat_dim, aggr=self.aggregation,
                                pre_layers=self.conv_pre_mlp_layers, post_layers=self.conv_post_mlp_layers,
                                use_edge_encoder=self.conv_use_edge_encoder)
            else:
                conv = RadarPointGNNConv(self.node_feat_dim, self.edge_feat_dim, aggr=self.aggregation,
                                         pre_layers=self.conv_pre_mlp_layers, post_layers=self.conv_post_mlp_layers)
            self.convs.append(conv)
            self.batch_norms.append(BatchNorm(next_layer_dim))
            layer_dim = next_layer_dim

        final_embedding_dim = self.conv_layer_dimensions[-1]

        layer_dimensions = config.classification_head_layer_dimensions[:-1]
        out_dim = config.classification_head_layer_dimensions[-1]
        self.classification_head = get_mlp(final_embedding_dim, out_dim, layer_dimensions, self.batch_norm_mlps)

        layer_dimensions = config.regression_head_layer_dimensions[:-1]
        out_dim = config.regression_head_layer_dimensions[-1]
        self.regression_head = get_mlp(final_embedding_dim, out_dim, layer_dimensions, self.batch_norm_mlps)

    def embed(self, x: torch.Tensor, edge_index: torch.Tensor, edge_attr: torch.Tensor) -> torch.Tensor:
        """The conv stack ``x = relu(batch_norm(conv(x, edge_index, edge_attr)))`` (gnn_models.py:124-128)
        after the optional embedding MLPs: the final node embeddings."""
        if self.initial_node_feature_embedding:
            x = _run_mlp(self.node_emb_mlp, x)
        if self.initial_edge_feature_embedding:
            edge_attr = _run_mlp(self.edge_emb_mlp, edge_attr)
        for conv, batch_norm in zip(self.convs, self.batch_norms):
            x = conv(x, edge_index, edge_attr)
            x = batch_norm(x, relu=True)
        return x

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor, edge_attr: torch.Tensor):
        """Returns ``(c, bb)``: class scores and regressed bounding box per node."""
        x = self.embed(x, edge_index, edge_attr)
        c = _run_mlp(self.classification_head, x)
        bb = _run_mlp(self.regression_head, x)
        return c, bb

    # ---- fused hot path: point cloud in, node embeddings out -----------------------------------
    def pipeline_config(self, graph_config) -> ops.PipelineConfig:
        """The conv stack + a GraphConstructionConfiguration as one fused-kernel configuration."""
        if self.initial_edge_feature_embedding:
            raise NotImplementedError("the one-call fused path feeds the conv stack with the raw edge attributes; with an "
                                      "edge embedding MLP forward_from_points runs the stage-by-stage CUDA path instead")
        layers = [conv.conv_params() for conv in self.convs]
        bn = [(b.module.weight, b.module.bias) for b in self.batch_norms]
        # The fused kernels normalise with BATCH statistics (the reference never leaves training mode,
        # SURVEY.md section 5).  A BatchNorm in eval mode would silently give other numbers than forward():
        if any(not b.module.training and b.module.running_mean is not None for b in self.batch_norms):
            raise RuntimeError("forward_from_points normalises with batch statistics (training-mode BatchNorm, as the "
                               "reference always runs it); call model.train() or use forward() for eval-mode BatchNorm")
        running = [(b.module.running_mean, b.module.running_var) for b in self.batch_norms]
        momenta = {0.1 if b.module.momentum is None else float(b.module.momentum) for b in self.batch_norms}
        if len(momenta) != 1:
            raise ValueError("the fused path takes one BatchNorm momentum for the whole stack")
        return ops.PipelineConfig(
            layers=layers, bn=bn, bn_running=running, bn_momentum=momenta.pop(),
            algorithm=graph_config.graph_construction_algorithm,
            k=graph_config.k if graph_config.k is not None else 6,
            r=graph_config.r if graph_config.r is not None else 1.0,
            distance_definition=graph_config.distance_definition, edge_features=list(graph_config.edge_features),
            edge_mode=graph_config.edge_mode, bn_eps=self.batch_norms[0].module.eps)

    def forward_from_points(self, graph_config, pos: torch.Tensor, vel: torch.Tensor, x: torch.Tensor,
                            frame_ptr=None, heads: bool = True):
        """Graph construction + conv stack in one fused call (rgnn_pipeline_forward), then the heads.
        ``pos`` / ``vel`` float32 ``[N, 2]``, ``x`` float32 ``[N, C0]``, ``frame_ptr`` ``[F + 1]`` point
        offsets of the frames of the batch.  Returns ``(edge_index, edge_attr, h)`` or, with heads,
        ``(edge_index, edge_attr, c, bb)``."""
        if pos.dtype == torch.float64 or vel.dtype == torch.float64:
            # input contract of the fused call: float32 coordinates (create_graph_data's pos / vel dtype,
            # dataset_creation.py:810-811); fp64 point clouds go through GraphConstructor, which searches in fp64
            if not (torch.equal(pos.float().double(), pos.double()) and torch.equal(vel.float().double(), vel.double())):
                raise ValueError("forward_from_points takes float32 coordinates; these float64 values are not "
                                 "float32-representable, so neighbours / edge_attr could differ from GraphConstructor's")
        if self.initial_node_feature_embedding:
            x = _run_mlp(self.node_emb_mlp, x)   # per node, independent of the graph: ahead of the fused call
        if self.initial_edge_feature_embedding:
            # the edge embedding MLP (gnn_models.py:120-121) sits between edge_attr and the conv stack: neighbour
            # search, edge features, embedding and conv stack as separate CUDA stages (same kernels, one call each)
            basis = pos if graph_config.distance_definition == "X" else torch.cat([pos, vel], dim=1)
            if graph_config.graph_construction_algorithm == "knn":
                edge_index = ops.knn_graph(basis, graph_config.k, frame_ptr)
            else:
                edge_index = ops.radius_graph(basis, graph_config.r, frame_ptr)
            edge_attr = ops.edge_features(pos, vel, edge_index, list(graph_config.edge_features), graph_config.edge_mode,
                                          out_dtype=torch.float32)
            ea = _run_mlp(self.edge_emb_mlp, edge_attr)
            h = x
            for conv, batch_norm in zip(self.convs, self.batch_norms):
                h = batch_norm(conv(h, edge_index, ea), relu=True)
            if not heads:
                return edge_index, edge_attr, h
            return edge_index, edge_attr, _run_mlp(self.classification_head, h), _run_mlp(self.regression_head, h)
        edge_index, edge_attr, h = ops.pipeline_forward(self.pipeline_config(graph_config), pos, vel, x, frame_ptr)
        for b in self.batch_norms:   # torch.nn.BatchNorm1d bookkeeping of a training-mode forward
            if b.module.num_batches_tracked is not None:
                b.module.num_batches_tracked += 1
        if not heads:
            return edge_index, edge_attr, h
        return edge_index, edge_attr, _run_mlp(self.classification_head, h), _run_mlp(self.regression_head, h)


def get_mlp(in_size: int, out_size: int, hidden_layer_sizes: List[int], batch_norm: bool) -> Sequential:
    """MLP with the specified layer number and dimension (reference gnn_models.py:137-178):
    Linear, then ([BatchNorm,] ReLU, Linear) per further layer."""
    if len(hidden_layer_sizes) == 0:
        modules = [Linear(in_size, out_size)]
    else:
        modules = [Linear(in_size, hidden_layer_sizes[0])]
        in_size = hidden_layer_sizes[0]

        if len(hidden_layer_sizes) == 1:
            layer_size = hidden_layer_sizes[0]
        else:
            for layer_size in hidden_layer_sizes[1:]:
                if batch_norm:
                    modules += [BatchNorm(in_size)]
                modules += [ReLU()]
                modules += [Linear(in_size, layer_size)]
                in_size = layer_size

        if batch_norm:
            modules += [BatchNorm(layer_size)]
        modules += [ReLU()]
        modules += [Linear(layer_size, out_size)]

    return Sequential(*modules)
