"""The reference's own unit tests for the hot path (test/test_graph_constructor.py,
test/test_preprocessor.py:207-257, test/test_gnn.py), re-run against the CUDA-backed mirror
classes, plus golden-vector checks of the full DetNetBasic forward."""
import numpy as np
import pytest
import torch

from helpers import DETNET_FIXTURES, GRAPH_FIXTURES, assert_edge_features_close, load_graph_fixture, load_module_fixture
from oracle import graph_oracle as go
from oracle import mpnn_oracle as mo

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def gr():
    from radargnn_b200.graph_constructor import graph
    return graph


@pytest.fixture(scope="module")
def gnn():
    import radargnn_b200.gnn as g
    return g


# ---- reference test/test_graph_constructor.py ---------------------------------------------------
def test_point_pair_features():
    from radargnn_b200.graph_constructor.features import get_En_equivariant_point_pair_metrics
    p1 = np.array([1, 1]).reshape(2, 1); p2 = np.array([3, 2]).reshape(2, 1)
    v1 = np.array([0, 1]).reshape(2, 1); v2 = np.array([1, 0]).reshape(2, 1)
    res = get_En_equivariant_point_pair_metrics(p1, p2, v1, v2, "directed")
    assert [round(v, 2) for v in res] == [2.24, 90.0, 63.43, 26.57]
    res = get_En_equivariant_point_pair_metrics(p1, p2, v1, np.zeros((2, 1)), "directed")
    assert [round(v, 2) for v in res] == [2.24, 90.0, 63.43, 90.0]


def test_edge_features(gr):
    X = np.array([[1, 1], [3, 2]]); V = np.array([[0, 1], [1, 0]])
    graph = gr.GeometricGraph()
    graph.X = X; graph.V = V; graph.F = {"rcs": np.array([0, 1]).reshape(2, 1)}
    graph.build(X, "knn", k=1)
    features = ["point_pair_features", "spatial_euclidean_distance", "velocity_euclidean_distance",
                "relative_position", "relative_velocity"]
    graph.extract_node_pair_features(features, "directed")
    assert [2.24, 90, 63.43, 26.57, 2.24, 1.41, -2, -1, -1, 1] == np.round(graph.E_feat[0, :], 2).tolist()


def test_node_features(gr):
    X = np.array([[1, 1], [3, 2]]); V = np.array([[0, 1], [1, 0]])
    graph = gr.GeometricGraph()
    graph.X = X; graph.V = V
    graph.F = {"rcs": np.array([1.8, 2.6]).reshape(2, 1), "time_index": np.array([100, 101]).reshape(2, 1)}
    graph.build(X, "knn", k=1)
    graph.extract_single_node_features(["rcs", "time_index", "degree", "velocity_vector_length",
                                        "velocity_vector", "spatial_coordinates"])
    assert [2.6, 101, 1, 1, 1, 0, 3, 2] == graph.X_feat[1, :].tolist()


def test_add_degree_to_inv_features(gr):
    graph = gr.GeometricGraph()
    graph.build(np.array([[1, 1], [3, 2]]), "knn", k=1)
    graph.add_degree_to_inv_features()
    graph.add_degree_to_inv_features()
    assert np.sum(graph.F.get("degree") == np.array([[1, 1], [1, 1]])) == 4
    assert graph.A.tolist() == [[0.0, 1.0], [1.0, 0.0]]      # lazily materialised adjacency


def test_add_node_feature(gr):
    X = np.array([[1, 1], [3, 2]]).reshape(2, 2); rcs = np.array([-1, -2]).reshape(2, 1)
    graph = gr.GeometricGraph()
    graph.X = X; graph.F = {"rcs": rcs}
    graph.build(X, "knn", k=1)
    graph.add_node_features(rcs)
    graph.extract_single_node_features(["rcs"])
    graph.add_node_features(rcs)
    assert (graph.X_feat[0, :] == [-1, -1, -1]).all()
    with pytest.raises(Exception, match="Feature dimension not compatible"):
        graph.add_node_features(np.zeros((3, 1)))


def test_build_ignores_unknown_routine_and_single_points(gr):
    graph = gr.Graph()
    graph.build(np.array([[1.0, 1.0]]), "knn", k=1)       # graph.py:45
    assert graph.E is None
    graph.build(np.array([[1.0, 1.0], [2.0, 2.0]]), "delaunay")   # graph.py:45-50
    assert graph.E is None
    with pytest.raises(ValueError):
        graph.build(np.array([[1.0, 1.0], [2.0, 2.0]]), "knn", k=2)


# ---- reference test/test_preprocessor.py:207-257 ----------------------------------------------------
def test_graph_constructor():
    from radargnn_b200.preprocessor import GraphConstructionConfiguration, GraphConstructor, RadarPointCloud
    pc = RadarPointCloud()
    pc.X_cc = np.array([[1, 1], [3, 2], [5, 8]]).reshape(3, 2)
    pc.V_cc_compensated = np.ones_like(pc.X_cc)
    pc.timestamp = np.array([100, 101, 102]).reshape(3, 1)
    config = GraphConstructionConfiguration("knn", {"k": 1, "r": 1}, ["spatial_coordinates", "time_index"],
                                            ["spatial_euclidean_distance"], "directed", "X")
    graph = GraphConstructor.build_geometric_graph(config, pc)
    assert (graph.E_feat[0, :] == 5 ** 0.5).all()
    assert (graph.X_feat[1, :] == np.array([3, 2, 1])).all()
    assert (graph.E == np.array([[0, 1], [1, 0], [2, 1]])).all()


def test_graph_constructor_distance_definition():
    from radargnn_b200.preprocessor import GraphConstructionConfiguration, GraphConstructor, RadarPointCloud
    pc = RadarPointCloud()
    pc.X_cc = np.array([[1, 1], [2, 2], [10, 10]]).reshape(3, 2)
    pc.V_cc_compensated = np.ones_like(pc.X_cc)
    pc.V_cc_compensated[0, :] = 100
    kw = (["spatial_coordinates"], ["spatial_euclidean_distance"], "directed")
    g = GraphConstructor.build_geometric_graph(GraphConstructionConfiguration("knn", {"k": 1, "r": 1}, *kw, "X"), pc)
    assert (g.E == np.array([[0, 1], [1, 0], [2, 1]])).all()
    g = GraphConstructor.build_geometric_graph(GraphConstructionConfiguration("knn", {"k": 1, "r": 1}, *kw, "XV"), pc)
    assert (g.E == np.array([[0, 1], [1, 2], [2, 1]])).all()
    with pytest.raises(Exception, match="Invalid graph construction algorithm selected"):
        GraphConstructionConfiguration("delaunay", {}, [], [], "directed", "X")


@pytest.mark.parametrize("name", GRAPH_FIXTURES)
def test_build_geometric_graph_matches_reference_golden(name):
    """Whole boundary function against vectors produced by the reference's own classes."""
    from radargnn_b200.preprocessor import (GraphConstructionConfiguration, GraphConstructor, RadarPointCloud,
                                            create_graph_data)
    fx = load_graph_fixture(name)
    pc = RadarPointCloud()
    pc.X_cc, pc.V_cc_compensated, pc.rcs, pc.timestamp = fx["X_cc"], fx["V_cc"], fx["rcs"], fx["timestamp"]
    config = GraphConstructionConfiguration(fx["algorithm"], {"k": fx["k"], "r": fx["r"]}, fx["node_features"],
                                            fx["edge_features"], fx["edge_mode"], fx["distance_definition"])
    graph = GraphConstructor.build_geometric_graph(config, pc)
    if fx["algorithm"] == "knn":
        np.testing.assert_array_equal(graph.E, fx["E"])
        assert_edge_features_close(graph.E_feat, fx["E_feat"], fx["edge_features"])
    else:
        E, EF = go.canonicalise_rows(fx["E"], fx["E_feat"])
        np.testing.assert_array_equal(graph.E, E)
        assert_edge_features_close(graph.E_feat, EF, fx["edge_features"])
    np.testing.assert_allclose(graph.X_feat, fx["X_feat"], rtol=1e-14, atol=0)
    data = create_graph_data(graph)
    assert data.x.dtype == torch.float32 and data.edge_index.dtype == torch.int64
    assert tuple(data.edge_index.shape) == (2, fx["E"].shape[0]) and data.edge_attr.dtype == torch.float32


@pytest.mark.parametrize("algo,dd,node_features,edge_features,edge_mode", [
    ("knn", "X", ["rcs", "spatial_coordinates"], ["relative_position"], "directed"),
    ("knn", "XV", ["rcs", "time_index", "degree", "velocity_vector_length", "velocity_vector"],
     ["point_pair_features", "relative_velocity"], "directed"),
    ("radius", "X", ["time_index", "degree"], ["spatial_euclidean_distance", "point_pair_features"], "undirected"),
])
def test_build_geometric_graphs_equals_per_frame_calls(algo, dd, node_features, edge_features, edge_mode):
    """Batched boundary function (all frames of a list in one device pass, the loop of dataset_creation.py:651-660):
    every graph equals the one ``build_geometric_graph`` builds for that frame alone, bit for bit."""
    from radargnn_b200 import synthetic
    from radargnn_b200.preprocessor import GraphConstructionConfiguration, GraphConstructor, RadarPointCloud
    clouds = []
    for s, n in enumerate([300, 7, 120, 2, 451]):
        fr = synthetic.radar_frame(n, seed=60 + s)
        pc = RadarPointCloud()
        pc.X_cc, pc.V_cc_compensated = fr.X_cc, fr.V_cc_compensated
        rng = np.random.default_rng(s)
        pc.rcs = rng.normal(size=(n, 1))
        pc.timestamp = rng.integers(0, 4, size=(n, 1)).astype(np.float64) * 1.7e4 + 1.6e15
        clouds.append(pc)
    config = GraphConstructionConfiguration(algo, {"k": 1 if algo == "knn" else 6, "r": 4.0}, node_features, edge_features,
                                            edge_mode, dd)
    got = GraphConstructor.build_geometric_graphs(config, clouds)
    assert len(got) == len(clouds)
    for g, pc in zip(got, clouds):
        want = GraphConstructor.build_geometric_graph(config, pc)
        np.testing.assert_array_equal(g.E, want.E)
        np.testing.assert_array_equal(g.E_feat, want.E_feat)
        np.testing.assert_array_equal(g.X_feat, want.X_feat)
        np.testing.assert_array_equal(g.A, want.A)
        assert list(g.F) == list(want.F)
        for name in g.F:
            np.testing.assert_array_equal(g.F[name], want.F[name])
        assert g.get_degree() == want.get_degree()
    # a frame the single-frame function rejects (k >= n) is rejected by the batch as well
    big_k = GraphConstructionConfiguration("knn", {"k": 7, "r": 1}, node_features, edge_features, edge_mode, dd)
    with pytest.raises(ValueError):
        GraphConstructor.build_geometric_graphs(big_k, clouds)
    assert GraphConstructor.build_geometric_graphs(config, []) == []


# ---- reference test/test_gnn.py ------------------------------------------------------------------------
def _ones(seq):
    for layer in seq:
        if hasattr(layer, "weight"):
            layer.weight = torch.nn.Parameter(torch.ones_like(layer.weight))
            layer.bias = torch.nn.Parameter(torch.zeros_like(layer.bias))


def test_get_mlp(gnn):
    mlp = gnn.get_mlp(2, 3, [5], False).to(DEV)
    _ones(mlp)
    x = torch.tensor([1, 1], dtype=torch.float32, device=DEV)
    assert mlp[0].weight.shape == (5, 2) and mlp[2].weight.shape == (3, 5)
    assert (mlp(x).detach().cpu().numpy() == np.array([10, 10, 10])).all()   # reference test_gnn.py:23 detaches too


def test_det_net_basic_layer_types(gnn):
    model = gnn.DetNetBasic(gnn.GNNArchitectureConfig(2, 3, [5], [3], [3], conv_layer_type="MPNNConv"))
    assert isinstance(model.convs[0], gnn.MPNNConv)
    model = gnn.DetNetBasic(gnn.GNNArchitectureConfig(2, 3, [2], [3], [3], conv_layer_type="RadarPointGNNConv"))
    assert isinstance(model.convs[0], gnn.RadarPointGNNConv)
    with pytest.raises(Exception, match="is invalid GNN conv layer type"):
        gnn.DetNetBasic(gnn.GNNArchitectureConfig(2, 3, [2], [3], [3], conv_layer_type="GATConv"))


def test_radar_point_gnn_conv(gnn):
    conv = gnn.RadarPointGNNConv(2, 1, "max", 2, 1).to(DEV)
    _ones(conv.pre_mlp); _ones(conv.post_mlp)
    assert len(conv.pre_mlp) == 3 and len(conv.post_mlp) == 1
    pre_x = torch.tensor([[1, 1, 1], [2, 2, 2]], dtype=torch.float32, device=DEV)
    from radargnn_b200.gnn.mpnn_layers import _run_sequential
    assert _run_sequential(conv.pre_mlp, pre_x).cpu().tolist() == [[9.0] * 3, [18.0] * 3]


def test_mpnn_conv_structure_and_forward(gnn):
    conv = gnn.MPNNConv(2, 4, 3, post_layers=2).to(DEV)
    _ones(conv.pre_mlp); _ones(conv.post_mlp)
    assert len(conv.pre_mlp) == 1 and len(conv.post_mlp) == 3
    from radargnn_b200.gnn.mpnn_layers import _run_sequential
    assert _run_sequential(conv.pre_mlp, torch.ones(1, 7, device=DEV))[0].cpu().tolist() == [7.0] * 7
    assert _run_sequential(conv.post_mlp, torch.full((1, 9), 2.0, device=DEV))[0].cpu().tolist() == [72.0] * 4
    # test/test_gnn.py:119-172: max aggregation over two parallel edges -> 436
    x = torch.tensor([[1, 1], [2, 2]], dtype=torch.float32, device=DEV)
    ei = torch.tensor([[0, 1, 0], [1, 0, 1]], device=DEV)
    ea = torch.tensor([[3, 3, 3], [4, 4, 4], [1, 1, 1]], dtype=torch.float32, device=DEV)
    out = conv.forward(x, ei, ea)
    assert (out[1, :].detach().cpu().numpy() == 436).all()
    # message() keeps the reference semantics: pre_mlp([x_i ; x_j ; e])
    m = conv.message(x[ei[1]], x[ei[0]], ea)
    assert m[0].cpu().tolist() == [15.0] * 7


def test_mpnn_conv_edge_encoder(gnn):
    conv = gnn.MPNNConv(1, 4, 2, use_edge_encoder=True).to(DEV)
    _ones(conv.pre_mlp); _ones(conv.post_mlp)
    conv.edge_encoder.weight = torch.nn.Parameter(torch.full_like(conv.edge_encoder.weight, 2.0))
    conv.edge_encoder.bias = torch.nn.Parameter(torch.zeros_like(conv.edge_encoder.bias))
    assert conv.pre_mlp[0].weight.shape[1] == 3
    x = torch.tensor([[1], [2]], dtype=torch.float32, device=DEV)
    ei = torch.tensor([[0, 1], [1, 0]], device=DEV)
    ea = torch.tensor([[1, 1], [2, 2]], dtype=torch.float32, device=DEV)
    assert conv.forward(x, ei, ea)[1, 0].item() == 23.0


# ---- DetNetBasic against vectors produced by the reference's own classes ---------------------------------
@pytest.mark.parametrize("name", DETNET_FIXTURES)
def test_detnet_matches_reference_golden(gnn, name):
    params, meta, d = load_module_fixture(name)
    cfg = gnn.GNNArchitectureConfig(
        int(meta["node_feature_dimension"]), int(meta["edge_feature_dimension"]),
        list(meta["conv_layer_dimensions"]), list(meta["classification_head_layer_dimensions"]),
        list(meta["regression_head_layer_dimensions"]),
        initial_node_feature_embedding=bool(meta.get("initial_node_feature_embedding", False)),
        initial_edge_feature_embedding=bool(meta.get("initial_edge_feature_embedding", False)),
        node_feature_embedding_layer_dimensions=meta.get("node_feature_embedding_layer_dimensions"),
        edge_feature_embedding_layer_dimensions=meta.get("edge_feature_embedding_layer_dimensions"),
        conv_layer_type=meta["conv_layer_type"], batch_norm_in_mlps=bool(meta.get("batch_norm_in_mlps", False)),
        conv_pre_mlp_layer_number=int(meta.get("conv_pre_mlp_layer_number", 1)),
        conv_post_mlp_layer_number=int(meta.get("conv_post_mlp_layer_number", 1)),
        aggregation_function=meta.get("aggregation_function", "max"))
    model = gnn.DetNetBasic(cfg)
    missing, unexpected = model.load_state_dict(params, strict=True), None   # the reference's key names load as is
    model = model.to(DEV)
    c, bb = model(d["x"].to(DEV), d["edge_index"].to(DEV), d["edge_attr"].to(DEV))
    assert mo.relative_error(c.cpu(), d["out_cls"]) <= 1e-4
    assert mo.relative_error(bb.cpu(), d["out_bb"]) <= 1e-4
    # training-mode BatchNorm updated its running statistics like torch's BatchNorm1d
    assert int(model.batch_norms[0].module.num_batches_tracked) == int(params["batch_norms.0.module.num_batches_tracked"]) + 1


def test_detnet_fused_path_equals_layerwise(gnn):
    from radargnn_b200 import synthetic
    from radargnn_b200.preprocessor import GraphConstructionConfiguration
    torch.manual_seed(0)
    cfg = gnn.GNNArchitectureConfig(8, 2, [16, 16], [4], [8, 5], batch_norm_in_mlps=False)
    model = gnn.DetNetBasic(cfg).to(DEV)
    gcfg = GraphConstructionConfiguration("knn", {"k": 8}, ["rcs"], ["relative_position"], "directed", "X")
    frames = [synthetic.radar_frame(200, seed=s) for s in range(3)]
    X, V, ptr = synthetic.frame_batch(frames)
    pos, vel = torch.tensor(X, dtype=torch.float32, device=DEV), torch.tensor(V, dtype=torch.float32, device=DEV)
    x = torch.randn(X.shape[0], 8, device=DEV)
    ei, ea, c, bb = model.forward_from_points(gcfg, pos, vel, x, ptr)
    np.testing.assert_array_equal(ei.cpu().numpy().T, go.batched_edges([f.X_cc for f in frames], "knn", k=8))
    # the fused path updated the running statistics like one training-mode forward() does
    rm1 = [b.module.running_mean.clone() for b in model.batch_norms]
    rv1 = [b.module.running_var.clone() for b in model.batch_norms]
    assert all(int(b.module.num_batches_tracked) == 1 for b in model.batch_norms)
    for b in model.batch_norms:
        b.module.reset_running_stats()
    c2, bb2 = model(x, ei, ea)
    assert mo.relative_error(c, c2) <= 1e-5 and mo.relative_error(bb, bb2) <= 1e-5
    for b, rm, rv in zip(model.batch_norms, rm1, rv1):
        torch.testing.assert_close(b.module.running_mean, rm, rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(b.module.running_var, rv, rtol=1e-4, atol=1e-6)
    # eval-mode BatchNorm would use other statistics than the fused kernels: refused, not silently different
    model.eval()
    with pytest.raises(RuntimeError, match="batch statistics"):
        model.forward_from_points(gcfg, pos, vel, x, ptr)
    model.train()
    with pytest.raises(ValueError, match="float32-representable"):
        model.forward_from_points(gcfg, pos.double() + 1e-9, vel.double(), x, ptr)


def test_detnet_with_embedding_mlps_from_points(gnn):
    """The shipped architecture shape (configurations/configuration_radarscenes.yml:18-42: node and edge embedding
    MLPs in front of the conv stack) from a point cloud: forward_from_points == graph ops + forward()."""
    from radargnn_b200 import ops, synthetic
    from radargnn_b200.preprocessor import GraphConstructionConfiguration
    torch.manual_seed(1)
    cfg = gnn.GNNArchitectureConfig(5, 4, [32, 32, 16], [8, 4], [8, 5], initial_node_feature_embedding=True,
                                    initial_edge_feature_embedding=True, node_feature_embedding_layer_dimensions=[8, 16],
                                    edge_feature_embedding_layer_dimensions=[4, 8, 16], batch_norm_in_mlps=False)
    model = gnn.DetNetBasic(cfg).to(DEV)
    gcfg = GraphConstructionConfiguration("knn", {"k": 6}, ["rcs"], ["point_pair_features"], "directed", "X")
    frames = [synthetic.radar_frame(150, seed=s) for s in range(2)]
    X, V, ptr = synthetic.frame_batch(frames)
    pos, vel = torch.tensor(X, dtype=torch.float32, device=DEV), torch.tensor(V, dtype=torch.float32, device=DEV)
    x = torch.randn(X.shape[0], 5, device=DEV)
    ei, ea, c, bb = model.forward_from_points(gcfg, pos, vel, x, ptr)
    np.testing.assert_array_equal(ei.cpu().numpy().T, go.batched_edges([f.X_cc for f in frames], "knn", k=6))
    assert ea.shape == (ei.shape[1], 4) and c.shape == (X.shape[0], 4) and bb.shape == (X.shape[0], 5)
    c2, bb2 = model(x, ei, ea)
    assert mo.relative_error(c, c2) <= 1e-5 and mo.relative_error(bb, bb2) <= 1e-5
    # ... and against plain torch modules holding the same parameters (fp64)
    want_c, want_bb = mo.det_net_forward({k: v.detach().cpu() for k, v in model.state_dict().items()}, x.cpu(), ei.cpu(), ea.cpu(),
                                         n_layers=3, node_embedding=True, edge_embedding=True, dtype=torch.float64)
    assert mo.relative_error(c.cpu(), want_c) <= 1e-4 and mo.relative_error(bb.cpu(), want_bb) <= 1e-4
    # node embedding only: the one-call fused path takes the embedded x
    cfg2 = gnn.GNNArchitectureConfig(5, 2, [64, 64], [4], [5], initial_node_feature_embedding=True,
                                     node_feature_embedding_layer_dimensions=[16, 64])
    model2 = gnn.DetNetBasic(cfg2).to(DEV)
    gcfg2 = GraphConstructionConfiguration("knn", {"k": 6}, ["rcs"], ["relative_position"], "directed", "X")
    ei2, ea2, c3, bb3 = model2.forward_from_points(gcfg2, pos, vel, x, ptr)
    for b in model2.batch_norms:
        b.module.reset_running_stats()
    c4, bb4 = model2(x, ei2, ea2)
    assert mo.relative_error(c3, c4) <= 1e-5 and mo.relative_error(bb3, bb4) <= 1e-5


def test_cpu_tensors_are_rejected(gnn):
    conv = gnn.MPNNConv(2, 4, 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        conv(torch.zeros(2, 2), torch.zeros(2, 1, dtype=torch.int64), torch.zeros(1, 3))
