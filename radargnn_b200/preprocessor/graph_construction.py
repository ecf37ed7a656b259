"""The boundary function of the graph half: ``GraphConstructor.build_geometric_graph``
(reference preprocessor/radarscenes/dataset_creation.py:187-229; nuScenes twin
preprocessor/nuscenes/conversion.py:70-109) and ``create_graph_data`` (dataset_creation.py:786-814),
on top of the CUDA-backed ``GeometricGraph``."""
from __future__ import annotations

from typing import List, Optional, Sequence

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


    @staticmethod
    def build_geometric_graphs(config: GraphConstructionConfiguration,
                               point_clouds: Sequence[RadarPointCloud]) -> List[GeometricGraph]:
        """The graphs ``build_geometric_graph`` returns for every point cloud of the list, from ONE pass over the
        device: the frames travel as one array + ``frame_ptr`` (edges never cross frames), so the neighbour search,
        the edge features, the degrees, the time ranks and the node features are one launch sequence for the whole
        list instead of one per frame -- the loop of the reference's dataset creation
        (preprocessor/radarscenes/dataset_creation.py:651-660, 699, fanned out over ray workers there).  Every
        returned graph has frame-local node ids and equals the single-frame result bit for bit."""
        from .. import ops
        clouds = list(point_clouds)
        sizes = [int(np.asarray(pc.X_cc).shape[0]) for pc in clouds]
        if (not clouds or min(sizes) < 2 or config.graph_construction_algorithm not in ("knn", "radius")
                or config.distance_definition not in ("X", "XV") or any(pc.rcs is None for pc in clouds)):
            # nothing to batch, or a case whose (error) behaviour is the single-frame function's
            return [GraphConstructor.build_geometric_graph(config, pc) for pc in clouds]
        if not torch.cuda.is_available():
            raise RuntimeError("radargnn_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        dev = torch.device("cuda", torch.cuda.current_device())
        node_ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        n = int(node_ptr[-1])
        X = np.concatenate([np.asarray(pc.X_cc, dtype=np.float64) for pc in clouds], axis=0)
        V = np.concatenate([np.asarray(pc.V_cc_compensated, dtype=np.float64) for pc in clouds], axis=0)

        batch = GeometricGraph()
        batch.X, batch.V = X, V
        batch.F = {"rcs": np.concatenate([np.asarray(pc.rcs).reshape(sz, -1) for pc, sz in zip(clouds, sizes)], axis=0)}
        t_idx = None
        if "time_index" in config.node_features:
            ts = np.concatenate([np.asarray(pc.timestamp).reshape(-1) for pc in clouds]).astype(np.float64)
            t_idx = ops.time_index(torch.from_numpy(ts).to(dev), node_ptr).cpu().numpy()
            batch.add_invariant_feature("time_index", t_idx.reshape(n, 1))

        basis = X if config.distance_definition == "X" else np.concatenate((X, V), axis=1)
        basis_dev = torch.as_tensor(np.ascontiguousarray(basis), device=dev)
        if config.graph_construction_algorithm == "knn":
            edge_index = ops.knn_graph(basis_dev, config.k, node_ptr)
        else:
            edge_index = ops.radius_graph(basis_dev, config.r, node_ptr)
        batch.adopt_edges(n, edge_index)
        batch.extract_node_pair_features(config.edge_features, config.edge_mode)
        batch.extract_single_node_features(config.node_features)

        # split: rows of E are grouped by query point, hence by frame.  Frame-local node ids are formed on the device
        # (one subtraction over the edge list); the per-frame arrays are row ranges (views) of the batch's arrays
        node_ptr_dev = torch.from_numpy(node_ptr).to(dev)
        edge_ptr_dev = torch.searchsorted(edge_index[0].contiguous(), node_ptr_dev)
        offsets = torch.repeat_interleave(node_ptr_dev[:-1], edge_ptr_dev[1:] - edge_ptr_dev[:-1])
        E = (edge_index - offsets.unsqueeze(0)).t().contiguous().cpu().numpy()
        edge_ptr = edge_ptr_dev.cpu().numpy()
        graphs = []
        for f, pc in enumerate(clouds):
            n0, n1, e0, e1 = int(node_ptr[f]), int(node_ptr[f + 1]), int(edge_ptr[f]), int(edge_ptr[f + 1])
            g = GeometricGraph()
            g.X, g.V = pc.X_cc, pc.V_cc_compensated
            g.F = {"rcs": pc.rcs}
            if t_idx is not None:
                ts = np.asarray(pc.timestamp)
                g.add_invariant_feature("time_index", t_idx[n0:n1].reshape(ts.shape).astype(ts.dtype))
            if "degree" in batch.F:
                g.add_invariant_feature("degree", batch.F["degree"][n0:n1])
            g.adopt_edges(n1 - n0, None, E[e0:e1])
            g.E_feat = None if batch.E_feat is None else batch.E_feat[e0:e1]
            g.X_feat = None if batch.X_feat is None else batch.X_feat[n0:n1]
            graphs.append(g)
        return graphs


def build_geometric_graph(config: GraphConstructionConfiguration, point_cloud: RadarPointCloud) -> GeometricGraph:
    """nuScenes twin of the boundary (reference preprocessor/nuscenes/conversion.py:70-109)."""
    return GraphConstructor.build_geometric_graph(config, point_cloud)


def build_geometric_graphs(config: GraphConstructionConfiguration,
                           point_clouds: Sequence[RadarPointCloud]) -> List[GeometricGraph]:
    """Batched twin (see ``GraphConstructor.build_geometric_graphs``)."""
    return GraphConstructor.build_geometric_graphs(config, point_clouds)


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
