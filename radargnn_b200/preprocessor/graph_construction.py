"""The boundary function of the graph half: ``GraphConstructor.build_geometric_graph``
(reference preprocessor/radarscenes/dataset_creation.py:187-229; nuScenes twin
preprocessor/nuscenes/conversion.py:70-109) and ``create_graph_data`` (dataset_creation.py:786-814),
on top of the CUDA-backed ``GeometricGraph``."""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from ..graph_constructor.graph import GeometricGraph
from .configs import GraphConstructionConfiguration
from .radar_point_cloud import RadarPointCloud


class GraphConstructor():

    @staticmethod
    def build_geometric_graph(config: GraphConstructionConfiguration,
                              point_cloud: RadarPointCloud) -> GeometricGraph:
        """Builds a graph from a point cloud based on the configuration."""
        if config.distance_definition == "X":
            distance_basis = point_cloud.X_cc
        elif config.distance_definition == "XV":
            distance_basis = np.concatenate((point_cloud.X_cc, point_cloud.V_cc_compensated), axis=1)
        else:
            raise UnboundLocalError("distance_definition must be 'X' or 'XV'")  # the reference falls through

        graph = GeometricGraph()
        graph.X = point_cloud.X_cc
        graph.V = point_cloud.V_cc_compensated
        graph.F = {"rcs": point_cloud.rcs}

        if "time_index" in config.node_features:
            # rank of every timestamp among the sorted distinct values (dataset_creation.py:214-223)
            # on the device (rgnn_time_index): sort + dense rank of the frame's timestamps; float64 holds the
            # datasets' integer microsecond stamps exactly (< 2^53)
            from .. import ops
            ts = np.asarray(point_cloud.timestamp)
            if not torch.cuda.is_available():
                raise RuntimeError("radargnn_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
            dev = torch.device("cuda", torch.cuda.current_device())
            t_idx = ops.time_index(torch.from_numpy(ts.reshape(-1).astype(np.float64)).to(dev)).cpu().numpy()
            graph.add_invariant_feature("time_index", t_idx.reshape(ts.shape).astype(ts.dtype))

        graph.build(distance_basis, config.graph_construction_algorithm, k=config.k, r=config.r)
        graph.extract_node_pair_features(config.edge_features, config.edge_mode)
        graph.extract_single_node_features(config.node_features)
        return graph


def build_geometric_graph(config: GraphConstructionConfiguration, point_cloud: RadarPointCloud) -> GeometricGraph:
    """nuScenes twin of the boundary (reference preprocessor/nuscenes/conversion.py:70-109)."""
    return GraphConstructor.build_geometric_graph(config, point_cloud)


class GraphData:
    """Minimal stand-in for ``torch_geometric.data.Data`` (absent here): the tensors the model reads."""

    def __init__(self, x, edge_index, edge_attr, y=None, pos=None, vel=None):
        self.x, self.edge_index, self.edge_attr, self.y, self.pos, self.vel = x, edge_index, edge_attr, y, pos, vel

    def to(self, device):
        for k, v in vars(self).items():
            if isinstance(v, torch.Tensor):
                setattr(self, k, v.to(device))
        return self


def collate_graph_data(graphs) -> GraphData:
    """PyG's disjoint-union collate of a list of device-resident ``GraphData`` (what the reference's
    ``DataLoader(graph_list, batch_size)`` does on the CPU for every step, utils/data_handling.py:25-30):
    concatenation is memory plumbing (torch.cat), the node-id offsets of edge_index and the ``batch`` vector come
    from ``rgnn_collate_offsets``.  Returns a GraphData with the extra attributes ``batch`` and ``ptr``."""
    from .. import ops
    if len(graphs) == 0:
        raise ValueError("collate_graph_data needs at least one graph")
    nodes = [int(g.x.shape[0]) for g in graphs]
    edges = [int(g.edge_index.shape[1]) for g in graphs]
    node_ptr = np.concatenate([[0], np.cumsum(nodes)]).astype(np.int64)
    edge_ptr = np.concatenate([[0], np.cumsum(edges)]).astype(np.int64)
    cat = lambda name: (None if getattr(graphs[0], name) is None else torch.cat([getattr(g, name) for g in graphs], dim=0))
    edge_index = torch.cat([g.edge_index for g in graphs], dim=1).contiguous()
    edge_index, batch = ops.collate_offsets(edge_index, edge_ptr, node_ptr)
    out = GraphData(cat("x"), edge_index, cat("edge_attr"), cat("y"), cat("pos"), cat("vel"))
    out.batch, out.ptr = batch, torch.from_numpy(node_ptr).to(batch.device)
    return out


def create_graph_data(graph: GeometricGraph, y: Optional[np.ndarray] = None) -> GraphData:
    """dtype / layout conversion to the model's input contract (dataset_creation.py:786-814):
    x float32 [N, Fn], edge_index int64 [2, E], edge_attr float32 [E, De], pos / vel float32 [N, 2]."""
    x = torch.tensor(graph.X_feat, dtype=torch.float32)
    edge_index = torch.tensor(np.asarray(graph.E).T.copy(), dtype=torch.int64)
    edge_attr = torch.tensor(graph.E_feat, dtype=torch.float32)
    pos = torch.tensor(np.asarray(graph.X), dtype=torch.float32)
    vel = torch.tensor(np.asarray(graph.V), dtype=torch.float32)
    yt = None if y is None else torch.tensor(y, dtype=torch.float32)
    return GraphData(x, edge_index, edge_attr, yt, pos, vel)
