"""GPU parity tests at the BASELINE.json configurations, full size, through the
fused C-ABI call: edge_index bit-exact against the reference's own sklearn call, edge_attr exact (or to
acos rounding for angles), node embeddings against the fp64 oracle with BOTH metrics -- the tensor-scale
max norm and the element-wise row-scaled error -- at the north star's 1e-4."""
import numpy as np
import pytest
import torch

from oracle import graph_oracle as go
from oracle import mpnn_oracle as mo
from radargnn_b200 import synthetic
from test_gpu_ops import DEV, _dev, _pipeline_cfg, _stack_params

pytestmark = pytest.mark.gpu

TOL = 1e-4   # BASELINE.json north_star: node embeddings <= 1e-4 relative (fp32)


@pytest.fixture(scope="module")
def ops():
    from radargnn_b200 import ops as _ops
    return _ops


def _assert_embeddings(h, want):
    rep = mo.parity_report(h.cpu(), want)
    assert rep["max_norm"] <= TOL, rep
    assert rep["rowwise"] <= TOL, rep      # every element within 1e-4 of its own row's scale


def test_config1_radius_single_frame_reference_node_features(ops):
    """Config 1: one RadarScenes-like frame (300 points), radius r = 3 m, node features
    [rcs, velocity_vector, time_index, degree] (Fn = 5) from GraphConstructor.build_geometric_graph, edge
    features [relative_position], directed; 1 x MPNNConv(5 -> 64, max) + BatchNorm + ReLU.  C = 5 is not a
    multiple of 4: the CUDA-core contraction path."""
    from radargnn_b200.preprocessor import GraphConstructionConfiguration, GraphConstructor, RadarPointCloud
    fr = synthetic.radar_frame(300, seed=0)
    feats_n = ["rcs", "velocity_vector", "time_index", "degree"]
    cfg_g = GraphConstructionConfiguration("radius", {"k": 6, "r": 3.0}, feats_n, ["relative_position"], "directed", "X")
    pc = RadarPointCloud()
    pc.X_cc, pc.V_cc_compensated, pc.rcs, pc.timestamp = fr.X_cc, fr.V_cc_compensated, fr.rcs, fr.timestamp
    graph = GraphConstructor.build_geometric_graph(cfg_g, pc)
    E_ref, Ef_ref, Xf_ref = go.build_geometric_graph(
        X_cc=fr.X_cc, V_cc=fr.V_cc_compensated, rcs=fr.rcs, timestamp=fr.timestamp, algorithm="radius", k=6, r=3.0,
        node_feature_names=feats_n, edge_feature_names=["relative_position"], edge_mode="directed",
        distance_definition="X")                                    # radius rows in canonical (ascending column) order
    np.testing.assert_array_equal(graph.E, E_ref)
    np.testing.assert_allclose(graph.X_feat, Xf_ref, rtol=1e-14, atol=0)
    x0 = graph.X_feat.astype(np.float32)
    assert x0.shape == (300, 5)
    params = _stack_params(1, 5, 64, 2, "MPNNConv", seed=0)
    cfg = _pipeline_cfg(ops, params, 1, "MPNNConv", "max", algorithm="radius", k=6, r=3.0, distance_definition="X",
                        edge_features=["relative_position"], edge_mode="directed")
    ei, ea, h = ops.pipeline_forward(cfg, _dev(fr.X_cc, torch.float32), _dev(fr.V_cc_compensated, torch.float32), _dev(x0))
    np.testing.assert_array_equal(ei.cpu().numpy().T, E_ref)                          # bit-exact
    np.testing.assert_array_equal(ea.cpu().numpy(), Ef_ref.astype(np.float32))
    want = mo.conv_stack_forward(params, torch.from_numpy(x0), torch.from_numpy(E_ref.T.copy()),
                                 torch.from_numpy(Ef_ref.astype(np.float32)), 1, "MPNNConv", "max", dtype=torch.float64)
    _assert_embeddings(h, want)
    # the same layer through the mirrored module API (conv -> BatchNorm -> ReLU, gnn_models.py:124-128)
    from radargnn_b200.gnn import MPNNConv
    from radargnn_b200.gnn._message_passing import BatchNorm
    conv, bn = MPNNConv(5, 64, 2, aggr="max").to(DEV), BatchNorm(64).to(DEV)
    conv.load_state_dict({k[len("convs.0."):]: v for k, v in params.items() if k.startswith("convs.0.")})
    bn.load_state_dict({"module.weight": params["batch_norms.0.module.weight"], "module.bias": params["batch_norms.0.module.bias"],
                        "module.running_mean": torch.zeros(64), "module.running_var": torch.ones(64),
                        "module.num_batches_tracked": torch.tensor(0)})
    h2 = bn(conv(_dev(x0), ei, ea), relu=True)
    _assert_embeddings(h2, want)


def test_config3_64_frames_k20_8_layers_d128(ops):
    """Config 3, complete: 64 RadarScenes-like frames x 300 points, k = 20, [relative_position],
    all EIGHT MPNNConv(128 -> 128) layers, each followed by BatchNorm(train) + ReLU."""
    frames = [synthetic.radar_frame(300, seed=s) for s in range(64)]
    X, V, ptr = synthetic.frame_batch(frames)
    n = X.shape[0]
    params = _stack_params(8, 128, 128, 2, "MPNNConv", seed=5)
    x0 = synthetic.node_embeddings(n, 128, seed=2)
    cfg = _pipeline_cfg(ops, params, 8, "MPNNConv", "max", algorithm="knn", k=20)
    ei, ea, h = ops.pipeline_forward(cfg, _dev(X, torch.float32), _dev(V, torch.float32), _dev(x0), ptr)
    E = go.batched_edges([f.X_cc for f in frames], "knn", k=20, backend="sklearn")
    np.testing.assert_array_equal(ei.cpu().numpy().T, E)
    ef = go.edge_features(X, V, E, ["relative_position"], "directed").astype(np.float32)
    np.testing.assert_array_equal(ea.cpu().numpy(), ef)
    want = mo.conv_stack_forward(params, torch.from_numpy(x0), torch.from_numpy(E.T.copy()), torch.from_numpy(ef),
                                 8, "MPNNConv", "max", dtype=torch.float64)
    _assert_embeddings(h, want)


def test_config4_64_frames_2000_points_point_pair_features(ops):
    """Config 4, one GPU's share: 64 nuScenes-like frames x 2 000 points (N = 128 000, E = 2.56 M), k = 20,
    rotation-invariant point-pair features (De = 4), 4 x MPNNConv(64 -> 64)."""
    frames = [synthetic.nuscenes_frame(2000, seed=s) for s in range(64)]
    X, V, ptr = synthetic.frame_batch(frames)
    n = X.shape[0]
    params = _stack_params(4, 64, 64, 4, "MPNNConv", seed=6)
    x0 = synthetic.node_embeddings(n, 64, seed=4)
    cfg = _pipeline_cfg(ops, params, 4, "MPNNConv", "max", algorithm="knn", k=20, edge_features=["point_pair_features"])
    ei, ea, h = ops.pipeline_forward(cfg, _dev(X, torch.float32), _dev(V, torch.float32), _dev(x0), ptr)
    E = go.batched_edges([f.X_cc for f in frames], "knn", k=20, backend="sklearn")
    np.testing.assert_array_equal(ei.cpu().numpy().T, E)
    ef = go.edge_features(X, V, E, ["point_pair_features"], "directed").astype(np.float32)
    np.testing.assert_allclose(ea.cpu().numpy(), ef, rtol=2e-7, atol=2e-5)   # angles: acos rounding, degrees
    # the embeddings are checked on the kernel's own edge_attr (an angle off by one fp32 ulp is a different input)
    want = mo.conv_stack_forward(params, torch.from_numpy(x0), torch.from_numpy(E.T.copy()), ea.cpu(),
                                 4, "MPNNConv", "max", dtype=torch.float64)
    _assert_embeddings(h, want)


def test_config5_one_rank_125k(ops):
    """Config 5, one rank's frame: 125 000 points, k = 16 (E = 2 M), 4 x MPNNConv(64 -> 64)."""
    fr = synthetic.uniform_square(125_000, seed=7)
    params = _stack_params(4, 64, 64, 2, "MPNNConv", seed=0)
    x0 = synthetic.node_embeddings(125_000, 64, seed=7)
    cfg = _pipeline_cfg(ops, params, 4, "MPNNConv", "max", algorithm="knn", k=16)
    ei, ea, h = ops.pipeline_forward(cfg, _dev(fr.X_cc, torch.float32), _dev(fr.V_cc_compensated, torch.float32), _dev(x0))
    E = go.knn_edges_sklearn(fr.X_cc, 16)
    assert go.kth_gap_is_tie_free(fr.X_cc, E, 16)
    np.testing.assert_array_equal(ei.cpu().numpy().T, E)
    ef = go.edge_features(fr.X_cc, fr.V_cc_compensated, E, ["relative_position"], "directed").astype(np.float32)
    np.testing.assert_array_equal(ea.cpu().numpy(), ef)
    want = mo.conv_stack_forward(params, torch.from_numpy(x0), torch.from_numpy(E.T.copy()), torch.from_numpy(ef),
                                 4, "MPNNConv", "max", dtype=torch.float64)
    _assert_embeddings(h, want)


def test_headline_frame_through_the_boundary_function():
    """The boundary function itself (GraphConstructor.build_geometric_graph, dataset_creation.py:187-229) on one
    frame of the headline size: 100 k points, k = 16, node features [rcs, velocity_vector_length, time_index,
    degree], relative_position.  Edges bit-exact against sklearn, features against vectorised numpy."""
    import scipy.sparse as sp
    from radargnn_b200.preprocessor import GraphConstructionConfiguration, GraphConstructor, RadarPointCloud
    n, k = 100_000, 16
    fr = synthetic.uniform_square(n, seed=0)
    rng = np.random.default_rng(0)
    pc = RadarPointCloud()
    pc.X_cc, pc.V_cc_compensated = fr.X_cc, fr.V_cc_compensated
    pc.rcs = rng.normal(size=(n, 1))
    pc.timestamp = rng.integers(0, 7, size=(n, 1)).astype(np.float64) * 1.7e4 + 1.6e15
    cfg = GraphConstructionConfiguration("knn", {"k": k, "r": 1.0}, ["rcs", "velocity_vector_length", "time_index", "degree"],
                                         ["relative_position"], "directed", "X")
    g = GraphConstructor.build_geometric_graph(cfg, pc)
    E = go.knn_edges_sklearn(fr.X_cc, k)
    np.testing.assert_array_equal(g.E, E)
    np.testing.assert_array_equal(g.E_feat, fr.X_cc[E[:, 0]] - fr.X_cc[E[:, 1]])
    A = sp.coo_matrix((np.ones(len(E)), (E[:, 0], E[:, 1])), shape=(n, n)).tocsr()
    degree = np.asarray(((A + A.T) > 0).sum(axis=1)).reshape(-1)
    t_idx = np.unique(pc.timestamp.reshape(-1), return_inverse=True)[1]
    want = np.stack([pc.rcs[:, 0], np.linalg.norm(fr.V_cc_compensated, axis=1), t_idx.astype(np.float64),
                     degree.astype(np.float64)], axis=1)
    np.testing.assert_allclose(g.X_feat, want, rtol=1e-14, atol=0)


def test_headline_100k_against_fp64_truth(ops):
    """The headline shape against the fp64 oracle (not the fp32 restatement), both metrics."""
    fr = synthetic.uniform_square(100_000, seed=0)
    params = _stack_params(4, 64, 64, 2, "MPNNConv", seed=0)
    x0 = synthetic.node_embeddings(100_000, 64)
    cfg = _pipeline_cfg(ops, params, 4, "MPNNConv", "max", algorithm="knn", k=16)
    ei, ea, h = ops.pipeline_forward(cfg, _dev(fr.X_cc, torch.float32), _dev(fr.V_cc_compensated, torch.float32), _dev(x0))
    E = ei.cpu().numpy().T     # bit-exactness of the graph is test_pipeline_headline_100k_against_reference_calls
    want = mo.conv_stack_forward(params, torch.from_numpy(x0), torch.from_numpy(E.T.copy()), ea.cpu(),
                                 4, "MPNNConv", "max", dtype=torch.float64)
    _assert_embeddings(h, want)


@pytest.mark.parametrize("aggr", ["max", "mean"])
def test_folded_target_term_under_cancellation(ops, aggr):
    """Adversarial case for the folded formulation: the kernels evaluate (W_x + W_m W_t) x + W_m M' with
    M' WITHOUT the target term W_t x_t (isolated nodes cancel it with M' = -W_t x_t).  Here W_t x_t is
    ~1e3 times larger than everything else and W_m is chosen so that the reference's post_mlp output nearly
    cancels it: any rounding of the folded product shows up relative to the SMALL result."""
    g = torch.Generator().manual_seed(21)
    n, c, de = 2000, 64, 2
    p = 2 * c + de
    x = torch.randn(n, c, generator=g)
    E = go.knn_edges_bruteforce(np.random.default_rng(3).uniform(0, 40, (n, 2)), 6)
    keep = E[:, 1] < n - 200                       # the last 200 nodes never receive a message: isolated
    ei = torch.from_numpy(E[keep].T.copy())
    ea = torch.randn(ei.shape[1], de, generator=g)
    w_pre = torch.randn(p, p, generator=g) / p ** 0.5
    w_pre[:, :c] *= 1000.0                         # huge target term W_t
    w_post = torch.randn(c, p + c, generator=g) / (p + c) ** 0.5
    # W_x := -W_m W_t + small: the folded weight W_x + W_m W_t is the small remainder of a large cancellation
    w_m = w_post[:, c:]
    w_post[:, :c] = -(w_m.double() @ w_pre[:, :c].double()).float() + 0.1 * torch.randn(c, c, generator=g)
    params = {"pre_mlp.0.weight": w_pre, "pre_mlp.0.bias": torch.randn(p, generator=g) * 0.1,
              "post_mlp.0.weight": w_post, "post_mlp.0.bias": torch.randn(c, generator=g) * 0.1}
    dev = {k: v.to(DEV) for k, v in params.items()}
    cp = ops.ConvParams("MPNNConv", c, c, de, aggr, [(dev["pre_mlp.0.weight"], dev["pre_mlp.0.bias"])],
                        [(dev["post_mlp.0.weight"], dev["post_mlp.0.bias"])])
    got = ops.conv_forward(cp, x.to(DEV), ops.csc_build(ei.to(DEV), n), ea.to(DEV)).cpu()
    want = mo.mpnn_conv_forward(params, x, ei, ea, aggr, dtype=torch.float64)
    # the yardstick is what fp32 arithmetic in the REFERENCE's own operation order achieves on this input
    ref32 = mo.mpnn_conv_forward(params, x, ei, ea, aggr, dtype=torch.float32)
    floor = mo.rowwise_relative_error(ref32, want)
    err = mo.rowwise_relative_error(got, want)
    assert err <= max(10 * floor, 1e-4), (err, floor)
    isolated = torch.arange(n - 200, n)
    exact = torch.nn.functional.linear(torch.cat([x[isolated], torch.zeros(200, p)], 1).double(), w_post.double(),
                                       params["post_mlp.0.bias"].double())
    assert mo.rowwise_relative_error(got[isolated], exact) <= max(10 * floor, 1e-4)


def test_out_of_range_edge_index_raises_like_pyg(ops):
    """an edge_index entry outside [0, N) must raise (PyG's gather raises an index error),
    never be used as an address."""
    ei = torch.tensor([[0, 1, 2, 7], [1, 2, 0, 1]], dtype=torch.int64, device=DEV)
    with pytest.raises(IndexError, match=r"outside \[0, N\)"):
        ops.csc_build(ei, 4)
    ei = torch.tensor([[0, 1, 2], [1, -1, 0]], dtype=torch.int64, device=DEV)
    with pytest.raises(IndexError):
        ops.csc_build(ei, 4)
    pos = torch.zeros(4, 2, dtype=torch.float64, device=DEV)
    with pytest.raises(IndexError):
        ops.edge_features(pos, pos, torch.tensor([[0, 9], [1, 2]], dtype=torch.int64, device=DEV), ["relative_position"], "directed")
    assert ops.csc_build(torch.tensor([[0, 1], [1, 3]], dtype=torch.int64, device=DEV), 4).n_edges == 2


def test_non_finite_coordinates_raise_like_sklearn(ops):
    X = np.random.default_rng(0).uniform(0, 10, (200, 2))
    X[17, 1] = np.nan
    with pytest.raises(ValueError, match="NaN or infinity"):
        ops.knn_graph(_dev(X), 4)
    with pytest.raises(ValueError, match="NaN or infinity"):
        ops.radius_graph(_dev(X), 1.0)
    X[17, 1] = np.inf
    with pytest.raises(ValueError, match="NaN or infinity"):
        ops.knn_graph(_dev(X.astype(np.float32)), 4)
    X[17, 1] = 3.0
    assert ops.knn_graph(_dev(X), 4).shape == (2, 800)
    # the fused path reports it through its error flag
    params = _stack_params(1, 8, 8, 2, "MPNNConv", seed=1)
    cfg = _pipeline_cfg(ops, params, 1, "MPNNConv", "max", algorithm="knn", k=4)
    X[3, 0] = np.nan
    with pytest.raises(ValueError, match="NaN or infinity"):
        ops.pipeline_forward(cfg, _dev(X, torch.float32), _dev(X, torch.float32), _dev(synthetic.node_embeddings(200, 8)))
