"""Generate tests/golden/*.npz by RUNNING THE REFERENCE'S OWN CODE in this container.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Usage (builder container
only; needs /root/reference):

    python -m oracle.make_golden

Graph fixtures call the reference's ``GeometricGraph`` (graph_constructor/graph.py)
in the order ``GraphConstructor.build_geometric_graph`` does
(preprocessor/radarscenes/dataset_creation.py:203-227 -- that module itself needs
ray / radar_scenes / torch_geometric at import time, so its 25 lines of
orchestration are replayed here on the reference's own classes).
MPNN fixtures instantiate the reference's ``MPNNConv`` / ``RadarPointGNNConv`` /
``DetNetBasic`` on top of oracle/pyg_shim.py.  Inputs are seeded; every fixture
stores inputs, parameters and the reference's outputs.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import reference_loader  # noqa: E402
from radargnn_b200 import synthetic  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

ALL_EDGE_FEATURES = ["point_pair_features", "spatial_euclidean_distance",
                     "velocity_euclidean_distance", "relative_position", "relative_velocity"]
ALL_NODE_FEATURES = ["rcs", "time_index", "degree", "velocity_vector_length",
                     "velocity_vector", "spatial_coordinates"]


def reference_graph(frame, algorithm, k, r, node_features, edge_features, edge_mode, distance_definition):
    gr = reference_loader.load("graph_constructor.graph")
    if distance_definition == "X":
        basis = frame.X_cc
    else:
        basis = np.concatenate((frame.X_cc, frame.V_cc_compensated), axis=1)
    graph = gr.GeometricGraph()
    graph.X = frame.X_cc
    graph.V = frame.V_cc_compensated
    graph.F = {"rcs": frame.rcs}
    if "time_index" in node_features:
        timestamps = np.unique(frame.timestamp)
        t_idx = np.zeros_like(frame.timestamp)
        for i, _ in enumerate(timestamps):
            t_idx[np.where(frame.timestamp == timestamps[i])[0]] = int(i)
        graph.add_invariant_feature("time_index", t_idx)
    graph.build(basis, algorithm, k=k, r=r)
    graph.extract_node_pair_features(edge_features, edge_mode)
    graph.extract_single_node_features(node_features)
    return graph


def graph_fixture(name, frame, algorithm, k, r, node_features, edge_features, edge_mode, dd):
    g = reference_graph(frame, algorithm, k, r, node_features, edge_features, edge_mode, dd)
    np.savez_compressed(
        os.path.join(GOLDEN_DIR, name + ".npz"),
        X_cc=frame.X_cc, V_cc=frame.V_cc_compensated, rcs=frame.rcs, timestamp=frame.timestamp,
        algorithm=algorithm, k=-1 if k is None else k, r=-1.0 if r is None else r,
        node_features=np.array(node_features), edge_features=np.array(edge_features),
        edge_mode=edge_mode, distance_definition=dd,
        E=g.E.astype(np.int64), E_feat=g.E_feat.astype(np.float64),
        X_feat=np.asarray(g.X_feat, dtype=np.float64))
    print(f"{name}: N={frame.n} E={g.E.shape[0]} De={g.E_feat.shape[1]} Fn={g.X_feat.shape[1]}")


def ppf_pairs_fixture():
    ft = reference_loader.load("graph_constructor.features")
    rng = np.random.default_rng(11)
    n = 64
    p1 = rng.uniform(-50, 50, (n, 2)); p2 = rng.uniform(-50, 50, (n, 2))
    v1 = rng.normal(0, 5, (n, 2)); v2 = rng.normal(0, 5, (n, 2))
    v1[0:8] = 0.0                       # zero first velocity
    v2[4:12] = 0.0                      # zero second / both
    p2[12:16] = p1[12:16]               # coincident points
    v2[16:20] = v1[16:20] * 2.0         # parallel
    v2[20:24] = -v1[20:24]              # anti-parallel
    p2[24:28] = p1[24:28] + v1[24:28]   # connection vector parallel to v1
    p1, p2, v1, v2 = (a.astype(np.float32).astype(np.float64) for a in (p1, p2, v1, v2))
    out = {}
    for mode in ("directed", "undirected"):
        res = np.empty((n, 4))
        for i in range(n):
            with np.errstate(invalid="ignore"):
                res[i] = ft.get_En_equivariant_point_pair_metrics(
                    p1[i].reshape(2, 1), p2[i].reshape(2, 1), v1[i].reshape(2, 1), v2[i].reshape(2, 1), mode)
        out[mode] = res
    np.savez_compressed(os.path.join(GOLDEN_DIR, "ppf_pairs.npz"), p1=p1, p2=p2, v1=v1, v2=v2,
                        directed=out["directed"], undirected=out["undirected"])
    print("ppf_pairs: 64 pairs, both modes")


def _random_graph(n, e, seed, isolated=2):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, n, (e,), generator=g)
    dst = torch.randint(0, n - isolated, (e,), generator=g)  # last `isolated` nodes get no message
    # a few parallel edges
    src[:3] = src[3:6]
    dst[:3] = dst[3:6]
    return torch.stack([src, dst]).long()


def _save_module_case(name, module, x, edge_index, edge_attr, out, meta):
    sd = {k: v.detach().numpy() for k, v in module.state_dict().items()}
    arrays = {"param::" + k: v for k, v in sd.items()}
    arrays.update(x=x.numpy(), edge_index=edge_index.numpy(), edge_attr=edge_attr.numpy())
    if isinstance(out, tuple):
        arrays.update(out_cls=out[0].detach().numpy(), out_bb=out[1].detach().numpy())
    else:
        arrays.update(out=out.detach().numpy())
    arrays.update({"meta::" + k: np.array(v) for k, v in meta.items()})
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **arrays)
    print(f"{name}: N={x.shape[0]} E={edge_index.shape[1]}")


def mpnn_fixtures():
    layers = reference_loader.load("gnn.mpnn_layers")
    models = reference_loader.load("gnn.gnn_models")
    cfgs = reference_loader.load("gnn.configs")
    cases = [
        ("mpnn_max", dict(in_channels=8, out_channels=16, edge_dim=2, aggr="max")),
        ("mpnn_add", dict(in_channels=8, out_channels=16, edge_dim=2, aggr="add")),
        ("mpnn_mean", dict(in_channels=6, out_channels=5, edge_dim=3, aggr="mean")),
        ("mpnn_deep", dict(in_channels=5, out_channels=7, edge_dim=2, aggr="max", pre_layers=3, post_layers=2)),
        ("mpnn_encoder", dict(in_channels=4, out_channels=9, edge_dim=3, aggr="max", use_edge_encoder=True)),
        ("mpnn_wide", dict(in_channels=64, out_channels=64, edge_dim=2, aggr="max")),
    ]
    for i, (name, kw) in enumerate(cases):
        torch.manual_seed(100 + i)
        conv = layers.MPNNConv(**kw)
        n, e = (200, 1500) if name == "mpnn_wide" else (30, 140)
        x = torch.randn(n, kw["in_channels"])
        ei = _random_graph(n, e, 200 + i)
        ea = torch.randn(e, kw["edge_dim"])
        out = conv.forward(x, ei, ea)
        _save_module_case(name, conv, x, ei, ea, out, dict(kind="MPNNConv", **kw))
    for i, (name, kw) in enumerate([
            ("rpgnn_max", dict(init_node_dim=8, init_edge_dim=2, aggr="max")),
            ("rpgnn_add_deep", dict(init_node_dim=6, init_edge_dim=4, aggr="add", pre_layers=2, post_layers=2))]):
        torch.manual_seed(300 + i)
        conv = layers.RadarPointGNNConv(**kw)
        n, e = 30, 140
        x = torch.randn(n, kw["init_node_dim"])
        ei = _random_graph(n, e, 400 + i)
        ea = torch.randn(e, kw["init_edge_dim"])
        out = conv.forward(x, ei, ea)
        _save_module_case(name, conv, x, ei, ea, out, dict(kind="RadarPointGNNConv", **kw))
    # full DetNetBasic (train-mode BatchNorm after every conv, heads, embeddings)
    det_cases = [
        ("detnet_mpnn", dict(node_feature_dimension=5, edge_feature_dimension=2, conv_layer_dimensions=[16, 16, 8],
                             classification_head_layer_dimensions=[6], regression_head_layer_dimensions=[16, 5],
                             initial_node_feature_embedding=True, initial_edge_feature_embedding=True,
                             node_feature_embedding_layer_dimensions=[8, 12], edge_feature_embedding_layer_dimensions=[4, 6],
                             conv_layer_type="MPNNConv", batch_norm_in_mlps=False)),
        ("detnet_rpgnn", dict(node_feature_dimension=6, edge_feature_dimension=2, conv_layer_dimensions=[6, 6],
                              classification_head_layer_dimensions=[6], regression_head_layer_dimensions=[5],
                              conv_layer_type="RadarPointGNNConv", batch_norm_in_mlps=False,
                              aggregation_function="add")),
    ]
    for i, (name, kw) in enumerate(det_cases):
        torch.manual_seed(500 + i)
        model = models.DetNetBasic(cfgs.GNNArchitectureConfig(**kw))
        n, e = 60, 400
        x = torch.randn(n, kw["node_feature_dimension"])
        ei = _random_graph(n, e, 600 + i)
        ea = torch.randn(e, kw["edge_feature_dimension"])
        with torch.no_grad():
            out = model(x, ei, ea)   # module is in training mode, as in the reference
        meta = {k: (v if not isinstance(v, list) else np.array(v)) for k, v in kw.items()}
        _save_module_case(name, model, x, ei, ea, out, dict(kind="DetNetBasic", **meta))


def detection_fixtures():
    """Loss: the arithmetic of the reference's gnn/trainer.py:184-216 executed as written (torch CrossEntropyLoss with
    class weights + the per-node Python loop over torch HuberLoss; oracle/detection_oracle.detection_loss is those
    lines).  NMS: torchvision.ops.nms itself, the function postprocessing.py:408 calls, after the reference's shift of
    negative coordinates (:400-404)."""
    import torch
    import torchvision
    from oracle import detection_oracle as do
    g = torch.Generator().manual_seed(11)
    n, k, nb = 600, 6, 5
    cls, bb = torch.randn(n, k, generator=g) * 2, torch.randn(n, nb, generator=g) * 3
    y = torch.cat([torch.randint(0, k, (n, 1), generator=g).float(), torch.randn(n, nb, generator=g) * 3], dim=1)
    w = torch.rand(k, generator=g) + 0.5
    loss, loss_cls, loss_bb, num_bb = do.detection_loss(cls, bb, y, w, bg_index=5, alpha=0.75, beta=1.25)
    np.savez(os.path.join(GOLDEN_DIR, "detection_loss.npz"), cls=cls.numpy(), bb=bb.numpy(), y=y.numpy(), weight=w.numpy(),
             bg_index=5, alpha=0.75, beta=1.25, loss=loss, loss_cls=loss_cls, loss_bb=loss_bb, num_bb=num_bb)
    xy = torch.rand(400, 2, generator=g) * 40 - 10
    wh = torch.rand(400, 2, generator=g) * 7 + 0.1
    boxes = torch.cat([xy, xy + wh], dim=1)
    scores = torch.rand(400, generator=g)
    shift = abs(float(boxes.min())) + 100 if float(boxes.min()) < 0 else 0
    keeps = {f"keep_{int(t * 100)}": torchvision.ops.nms((boxes + shift).float(), scores.float(), t).numpy() for t in (0.1, 0.3, 0.6)}
    np.savez(os.path.join(GOLDEN_DIR, "nms_aligned.npz"), boxes=boxes.numpy(), scores=scores.numpy(), **keeps)


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    small = synthetic.radar_frame(n=48, seed=1, extent=(30.0, 30.0))
    graph_fixture("graph_knn3_directed_all", small, "knn", 3, None, ALL_NODE_FEATURES, ALL_EDGE_FEATURES, "directed", "X")
    graph_fixture("graph_knn5_undirected_all", small, "knn", 5, None, ALL_NODE_FEATURES,
                  ["spatial_euclidean_distance", "velocity_euclidean_distance", "relative_position",
                   "relative_velocity", "point_pair_features"], "undirected", "X")
    graph_fixture("graph_knn4_xv", small, "knn", 4, None, ["rcs", "velocity_vector", "time_index", "degree"],
                  ["relative_position"], "directed", "XV")
    frame = synthetic.radar_frame(n=300, seed=0)
    graph_fixture("graph_radius3_config1", frame, "radius", None, 3.0,
                  ["rcs", "velocity_vector", "time_index", "degree"], ["relative_position"], "directed", "X")
    graph_fixture("graph_knn20_shipped", frame, "knn", 20, None,
                  ["rcs", "velocity_vector", "time_index", "degree"], ["relative_position"], "directed", "X")
    uni = synthetic.uniform_square(400, seed=3)
    graph_fixture("graph_knn16_uniform", uni, "knn", 16, None, ["spatial_coordinates"],
                  ["relative_position", "point_pair_features"], "directed", "X")
    tiny = synthetic.radar_frame(n=9, seed=5, extent=(10.0, 10.0))   # sklearn brute-force branch (k >= N//2)
    graph_fixture("graph_knn4_tiny_brute", tiny, "knn", 4, None, ["degree"], ["spatial_euclidean_distance"], "directed", "X")
    graph_fixture("graph_radius4_tiny_brute", tiny, "radius", None, 4.0, ["degree"], ["relative_velocity"], "undirected", "X")
    ppf_pairs_fixture()
    mpnn_fixtures()
    detection_fixtures()


if __name__ == "__main__":
    main()
