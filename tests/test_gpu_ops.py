"""GPU parity tests of the C-ABI kernels (through radargnn_b200.ops) against the CPU oracle and
the committed golden vectors.  Bit-exact for edge_index / degree / CSC; fp tolerances stated
per test (node embeddings: <= 1e-4 relative, BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from helpers import (CONV_FIXTURES, DETNET_FIXTURES, GRAPH_FIXTURES, assert_edge_features_close,
                     load_graph_fixture, load_module_fixture)
from oracle import graph_oracle as go
from oracle import mpnn_oracle as mo
from radargnn_b200 import synthetic

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from radargnn_b200 import ops as _ops
    return _ops


def _dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV)


def _basis(fx):
    if fx["distance_definition"] == "XV":
        return np.concatenate([fx["X_cc"], fx["V_cc"]], axis=1)
    return fx["X_cc"]


# ---- neighbour search ---------------------------------------------------------------------
@pytest.mark.parametrize("name", GRAPH_FIXTURES)
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_edge_index_matches_reference_golden(ops, name, dtype):
    fx = load_graph_fixture(name)
    basis = _basis(fx)
    if dtype == torch.float32 and not np.array_equal(basis.astype(np.float32).astype(np.float64), basis):
        pytest.skip("fixture coordinates are not float32-representable")
    b = _dev(basis, dtype)
    if fx["algorithm"] == "knn":
        ei = ops.knn_graph(b, fx["k"]).cpu().numpy()
        np.testing.assert_array_equal(ei.T, fx["E"])          # bit-exact, including row order
    else:
        ei = ops.radius_graph(b, fx["r"]).cpu().numpy()
        np.testing.assert_array_equal(ei.T, go.canonicalise_rows(fx["E"]))  # canonical column order


@pytest.mark.parametrize("n,k,dims", [(2, 1, 2), (17, 16, 2), (300, 16, 2), (300, 20, 4), (5000, 16, 2),
                                      (1000, 33, 2), (1000, 64, 2), (3000, 3, 4)])
def test_knn_random_vs_oracle(ops, n, k, dims):
    rng = np.random.default_rng(n * 131 + k)
    X = rng.uniform(-50, 50, (n, dims)).astype(np.float32).astype(np.float64)
    want = go.knn_edges_bruteforce(X, k)
    for dt in (torch.float32, torch.float64):
        got = ops.knn_graph(_dev(X, dt), k).cpu().numpy().T
        np.testing.assert_array_equal(got, want)


def test_knn_clustered_and_degenerate_layouts(ops):
    fr = synthetic.radar_frame(600, seed=5)
    np.testing.assert_array_equal(ops.knn_graph(_dev(fr.X_cc), 20).cpu().numpy().T,
                                  go.knn_edges_bruteforce(fr.X_cc, 20))
    # all points on a line / duplicates: the documented tie rule is ascending (d2, index)
    X = np.stack([np.arange(40.0), np.zeros(40)], axis=1)
    np.testing.assert_array_equal(ops.knn_graph(_dev(X), 4).cpu().numpy().T, go.knn_edges_bruteforce(X, 4))
    X = np.zeros((9, 2)); X[5:] = 1.0
    np.testing.assert_array_equal(ops.knn_graph(_dev(X), 3).cpu().numpy().T, go.knn_edges_bruteforce(X, 3))


def test_knn_k_not_smaller_than_n_raises(ops):
    X = _dev(np.random.default_rng(0).uniform(0, 1, (5, 2)))
    with pytest.raises(ValueError, match="Expected n_neighbors < n_samples_fit"):
        ops.knn_graph(X, 5)


def test_batched_frames_disjoint_union(ops):
    frames = [synthetic.radar_frame(n, seed=s) for s, n in enumerate([300, 1, 57, 0, 120, 2])]
    X, V, ptr = synthetic.frame_batch(frames)
    want = go.batched_edges([f.X_cc for f in frames], "knn", k=1)
    got = ops.knn_graph(_dev(X), 1, ptr).cpu().numpy().T
    np.testing.assert_array_equal(got, want)
    frames = [f for f in frames if f.n != 2]
    X, V, ptr = synthetic.frame_batch(frames)
    want = go.batched_edges([f.X_cc for f in frames if True], "radius", r=4.0)
    got = ops.radius_graph(_dev(X), 4.0, ptr).cpu().numpy().T
    np.testing.assert_array_equal(got, want)
    big = [synthetic.radar_frame(300, seed=s) for s in range(16)]
    X, V, ptr = synthetic.frame_batch(big)
    np.testing.assert_array_equal(ops.knn_graph(_dev(X, torch.float32), 20, ptr).cpu().numpy().T,
                                  go.batched_edges([f.X_cc for f in big], "knn", k=20))


@pytest.mark.parametrize("n,r", [(2, 5.0), (300, 3.0), (300, 0.0), (2000, 2.5), (500, 1e9)])
def test_radius_random_vs_oracle(ops, n, r):
    fr = synthetic.radar_frame(n, seed=n)
    want = go.radius_edges_bruteforce(fr.X_cc, r)
    got = ops.radius_graph(_dev(fr.X_cc), r).cpu().numpy().T
    np.testing.assert_array_equal(got, want)


def test_large_knn_properties(ops):
    """Headline size (100 k points, k = 16): properties that do not need the O(N^2) oracle."""
    fr = synthetic.uniform_square(100_000, seed=0)
    X = _dev(fr.X_cc, torch.float32)
    ei = ops.knn_graph(X, 16)
    assert ei.shape == (2, 1_600_000)
    src, dst = ei[0].view(-1, 16), ei[1].view(-1, 16)
    assert torch.equal(src[:, 0], torch.arange(100_000, device=DEV))
    assert bool((src == src[:, :1]).all()) and bool((dst != src).all())
    Xd = X.double()
    d2 = ((Xd[src] - Xd[dst]) ** 2).sum(-1)
    assert bool((d2[:, 1:] >= d2[:, :-1]).all())                       # ascending distances
    assert bool((torch.sort(dst, dim=1).values.diff(dim=1) > 0).all())  # no duplicate neighbours
    # spot-check 256 random rows against a brute-force search on the device
    rows = torch.randint(0, 100_000, (256,), device=DEV)
    full = ((Xd[rows][:, None, :] - Xd[None, :, :]) ** 2)
    full = full[..., 0] + full[..., 1]
    full[torch.arange(256, device=DEV), rows] = float("inf")
    assert torch.equal(torch.topk(full, 16, dim=1, largest=False, sorted=True).values, d2[rows])
    # idempotence: a second build gives the same bytes
    assert torch.equal(ei, ops.knn_graph(X, 16))


# ---- edge / node features --------------------------------------------------------------------
@pytest.mark.parametrize("name", GRAPH_FIXTURES)
def test_edge_and_node_features_match_reference_golden(ops, name):
    fx = load_graph_fixture(name)
    E = fx["E"]
    ei = _dev(E.T.copy())
    got = ops.edge_features(_dev(fx["X_cc"]), _dev(fx["V_cc"]), ei, fx["edge_features"], fx["edge_mode"],
                            out_dtype=torch.float64).cpu().numpy()
    assert_edge_features_close(got, fx["E_feat"], fx["edge_features"])
    n = fx["X_cc"].shape[0]
    deg = ops.undirected_degree(_dev(go.canonicalise_rows(E).T.copy()), n)
    np.testing.assert_array_equal(deg.cpu().numpy(), go.undirected_degree(E, n))
    ti = go.time_index(fx["timestamp"])
    xf = ops.node_features(fx["node_features"], n, DEV, rcs=_dev(fx["rcs"]), time_index=_dev(ti), degree=deg,
                           pos=_dev(fx["X_cc"]), vel=_dev(fx["V_cc"])).cpu().numpy()
    np.testing.assert_allclose(xf, fx["X_feat"], rtol=1e-14, atol=0)


def test_edge_features_fp32_output_and_known_answers(ops):
    X = np.array([[1., 1.], [3., 2.]]); V = np.array([[0., 1.], [1., 0.]])
    ei = ops.knn_graph(_dev(X), 1)
    feats = ["point_pair_features", "spatial_euclidean_distance", "velocity_euclidean_distance",
             "relative_position", "relative_velocity"]
    ef = ops.edge_features(_dev(X), _dev(V), ei, feats, "directed", out_dtype=torch.float64).cpu().numpy()
    # reference test/test_graph_constructor.py:34-59
    assert np.round(ef[0], 2).tolist() == [2.24, 90, 63.43, 26.57, 2.24, 1.41, -2, -1, -1, 1]
    assert np.round(ef[1], 2).tolist() == [2.24, 90, 153.43, 116.57, 2.24, 1.41, 2, 1, 1, -1]
    # zero velocity -> 90 degrees (test_graph_constructor.py:20-31)
    V0 = np.array([[0., 1.], [0., 0.]])
    ef = ops.edge_features(_dev(X), _dev(V0), ei, ["point_pair_features"], "directed", torch.float64).cpu().numpy()
    assert np.round(ef[0], 2).tolist() == [2.24, 90.0, 63.43, 90.0]
    with pytest.raises(Exception, match="Invalid feature specified"):
        ops.edge_features(_dev(X), _dev(V), ei, ["no_such_feature"], "directed")
    # float32 inputs / outputs equal the fp64 computation rounded once
    fr = synthetic.nuscenes_frame(500, seed=3)
    E = go.knn_edges_bruteforce(fr.X_cc, 8)
    want = go.edge_features(fr.X_cc, fr.V_cc_compensated, E, feats, "undirected").astype(np.float32)
    got = ops.edge_features(_dev(fr.X_cc, torch.float32), _dev(fr.V_cc_compensated, torch.float32), _dev(E.T.copy()),
                            feats, "undirected").cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=2e-7, atol=2e-5)


# ---- CSC -----------------------------------------------------------------------------------
def test_csc_build_is_stable_counting_sort(ops):
    rng = np.random.default_rng(4)
    n, e = 500, 6000
    ei = rng.integers(0, n - 7, (2, e))
    csc = ops.csc_build(_dev(ei), n)
    order = np.argsort(ei[1], kind="stable")
    np.testing.assert_array_equal(csc.eid.cpu().numpy(), order)
    np.testing.assert_array_equal(csc.src.cpu().numpy(), ei[0][order])
    ptr = np.zeros(n + 1, dtype=np.int64); np.add.at(ptr, ei[1] + 1, 1)
    np.testing.assert_array_equal(csc.ptr.cpu().numpy(), np.cumsum(ptr))


# ---- layers -----------------------------------------------------------------------------------
def _conv_params(ops, params, meta, prefix=""):
    def seq(name):
        idx = mo._mlp_keys(params, prefix + name)
        return [(params[f"{prefix}{name}.{i}.weight"].to(DEV), params[f"{prefix}{name}.{i}.bias"].to(DEV)) for i in idx]
    pre, post = seq("pre_mlp"), seq("post_mlp")
    enc = None
    if prefix + "edge_encoder.weight" in params:
        enc = (params[prefix + "edge_encoder.weight"].to(DEV), params[prefix + "edge_encoder.bias"].to(DEV))
    kind = meta["kind"] if "kind" in meta and meta["kind"] in ("MPNNConv", "RadarPointGNNConv") else meta["conv_layer_type"]
    p = pre[0][0].shape[0]
    if kind == "MPNNConv":
        c = post[0][0].shape[1] - p
        de = enc[0].shape[1] if enc is not None else p - 2 * c
    else:
        c = post[0][0].shape[1] - p
        de = p - c
    return ops.ConvParams(kind, c, post[-1][0].shape[0], de, meta.get("aggr", meta.get("aggregation_function", "max")),
                          pre, post, enc)


@pytest.mark.parametrize("name", CONV_FIXTURES)
def test_conv_matches_reference_golden(ops, name):
    params, meta, d = load_module_fixture(name)
    cp = _conv_params(ops, params, meta)
    n = d["x"].shape[0]
    csc = ops.csc_build(d["edge_index"].to(DEV), n)
    out = ops.conv_forward(cp, d["x"].to(DEV), csc, d["edge_attr"].to(DEV)).cpu()
    assert out.shape == d["out"].shape
    assert mo.relative_error(out, d["out"]) <= 1e-5   # north star: <= 1e-4 relative (fp32)


def test_conv_known_answers(ops):
    # reference test/test_gnn.py:119-172 -> 436 and :175-221 -> 23
    ones = lambda *s: torch.ones(*s, device=DEV)
    zeros = lambda *s: torch.zeros(*s, device=DEV)
    cp = ops.ConvParams("MPNNConv", 2, 4, 3, "max", [(ones(7, 7), zeros(7))],
                        [(ones(4, 9), zeros(4)), (ones(4, 4), zeros(4))])
    x = torch.tensor([[1., 1.], [2., 2.]], device=DEV)
    ei = torch.tensor([[0, 1, 0], [1, 0, 1]], device=DEV)
    ea = torch.tensor([[3., 3., 3.], [4., 4., 4.], [1., 1., 1.]], device=DEV)
    out = ops.conv_forward(cp, x, ops.csc_build(ei, 2), ea)
    assert out[1].tolist() == [436.0] * 4
    cp = ops.ConvParams("MPNNConv", 1, 4, 2, "max", [(ones(3, 3), zeros(3))], [(ones(4, 4), zeros(4))],
                        (torch.full((1, 2), 2.0, device=DEV), zeros(1)))
    x = torch.tensor([[1.], [2.]], device=DEV)
    ei = torch.tensor([[0, 1], [1, 0]], device=DEV)
    ea = torch.tensor([[1., 1.], [2., 2.]], device=DEV)
    out = ops.conv_forward(cp, x, ops.csc_build(ei, 2), ea)
    assert out[1, 0].item() == 23.0


@pytest.mark.parametrize("kind,aggr,pre,post,c,cout,de", [
    ("MPNNConv", "max", 1, 1, 64, 64, 2), ("MPNNConv", "add", 1, 2, 16, 24, 4), ("MPNNConv", "mean", 2, 1, 8, 8, 2),
    ("MPNNConv", "min", 3, 2, 5, 7, 3), ("RadarPointGNNConv", "max", 1, 1, 32, 32, 2),
    ("RadarPointGNNConv", "add", 2, 2, 12, 12, 4), ("MPNNConv", "max", 1, 1, 128, 128, 2),
    # split message layout (C = 64 MPNNConv on the tensor-core path): every aggregation and tail width
    ("MPNNConv", "add", 1, 1, 64, 64, 4), ("MPNNConv", "mean", 1, 2, 64, 48, 3), ("MPNNConv", "min", 1, 1, 64, 32, 1),
])
def test_conv_random_vs_oracle(ops, kind, aggr, pre, post, c, cout, de):
    g = torch.Generator().manual_seed(c * 7 + de)
    n, e = 700, 9000
    x = torch.randn(n, c, generator=g)
    ei = torch.randint(0, n - 5, (2, e), generator=g)   # last nodes have zero in-degree
    ea = torch.randn(e, de, generator=g)
    p = 2 * c + de if kind == "MPNNConv" else c + de
    params = {}
    def lin(key, o, i):
        params[key + ".weight"] = torch.randn(o, i, generator=g) / i ** 0.5
        params[key + ".bias"] = torch.randn(o, generator=g) * 0.1
    for l in range(pre):
        lin(f"pre_mlp.{2 * l}", p, p)
    lin("post_mlp.0", cout, p + c)
    for l in range(1, post):
        lin(f"post_mlp.{2 * l}", cout, cout)
    if kind == "MPNNConv":
        want = mo.mpnn_conv_forward(params, x, ei, ea, aggr, dtype=torch.float64)
    else:
        want = mo.radar_point_gnn_conv_forward(params, x, ei, ea, aggr, dtype=torch.float64)
    cp = _conv_params(ops, params, {"kind": kind, "aggr": aggr})
    got = ops.conv_forward(cp, x.to(DEV), ops.csc_build(ei.to(DEV), n), ea.to(DEV)).cpu()
    assert mo.relative_error(got, want) <= 2e-5
    assert torch.isfinite(got).all()


@pytest.mark.parametrize("aggr", ["max", "add", "mean"])
def test_conv_split_layout_high_in_degree(ops, aggr):
    """Warp-per-node aggregate: in-degrees far above 32 (several slot batches per node), a node with exactly
    32 and 33 incoming edges, and nodes without any."""
    g = torch.Generator().manual_seed(5)
    n, c, de = 150, 64, 2
    dst = torch.cat([torch.randint(0, 100, (12000,), generator=g), torch.full((32,), 100), torch.full((33,), 101)])
    src = torch.randint(0, n, (dst.numel(),), generator=g)
    ei = torch.stack([src, dst])
    x = torch.randn(n, c, generator=g)
    ea = torch.randn(ei.shape[1], de, generator=g)
    p = 2 * c + de
    params = {"pre_mlp.0.weight": torch.randn(p, p, generator=g) / p ** 0.5, "pre_mlp.0.bias": torch.randn(p, generator=g) * 0.1,
              "post_mlp.0.weight": torch.randn(c, p + c, generator=g) / (p + c) ** 0.5, "post_mlp.0.bias": torch.randn(c, generator=g) * 0.1}
    want = mo.mpnn_conv_forward(params, x, ei, ea, aggr, dtype=torch.float64)
    cp = _conv_params(ops, params, {"kind": "MPNNConv", "aggr": aggr})
    got = ops.conv_forward(cp, x.to(DEV), ops.csc_build(ei.to(DEV), n), ea.to(DEV)).cpu()
    assert mo.relative_error(got, want) <= 2e-5
    assert torch.isfinite(got).all()


def test_batchnorm_relu_matches_torch(ops):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3000, 70, generator=g) * 3 + 1.5
    bn = torch.nn.BatchNorm1d(70)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5, generator=g); bn.bias.uniform_(-1, 1, generator=g)
    want = torch.relu(bn(x)).detach()
    rm, rv = torch.zeros(70, device=DEV), torch.ones(70, device=DEV)
    got = ops.batchnorm_relu(x.to(DEV), bn.weight.to(DEV), bn.bias.to(DEV), running_mean=rm, running_var=rv).cpu()
    assert mo.relative_error(got, want) <= 2e-6
    torch.testing.assert_close(rm.cpu(), bn.running_mean, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(rv.cpu(), bn.running_var, rtol=1e-5, atol=1e-6)


def test_linear_matches_torch(ops):
    g = torch.Generator().manual_seed(2)
    x, w, b = torch.randn(1000, 37, generator=g), torch.randn(29, 37, generator=g), torch.randn(29, generator=g)
    got = ops.linear(x.to(DEV), w.to(DEV), b.to(DEV), relu_input=True).cpu()
    want = torch.nn.functional.linear(torch.relu(x).double(), w.double(), b.double())
    assert mo.relative_error(got, want) <= 2e-6


# ---- fused pipeline ----------------------------------------------------------------------------
def _stack_params(n_layers, c0, c, de, kind, seed):
    g = torch.Generator().manual_seed(seed)
    params = {}
    def lin(key, o, i):
        bound = 1.0 / i ** 0.5
        params[key + ".weight"] = (torch.rand(o, i, generator=g) * 2 - 1) * bound
        params[key + ".bias"] = (torch.rand(o, generator=g) * 2 - 1) * bound
    cin = c0
    for l in range(n_layers):
        p = 2 * cin + de if kind == "MPNNConv" else cin + de
        lin(f"convs.{l}.pre_mlp.0", p, p)
        lin(f"convs.{l}.post_mlp.0", c if kind == "MPNNConv" else cin, p + cin)
        cout = c if kind == "MPNNConv" else cin
        params[f"batch_norms.{l}.module.weight"] = torch.rand(cout, generator=g) + 0.5
        params[f"batch_norms.{l}.module.bias"] = torch.rand(cout, generator=g) - 0.5
        cin = cout
    return params


def _pipeline_cfg(ops, params, n_layers, kind, aggr, **kw):
    layers, bn = [], []
    for l in range(n_layers):
        sp = mo.sub_params(params, f"convs.{l}")
        layers.append(_conv_params(ops, sp, {"kind": kind, "aggr": aggr}))
        bn.append((params[f"batch_norms.{l}.module.weight"].to(DEV), params[f"batch_norms.{l}.module.bias"].to(DEV)))
    return ops.PipelineConfig(layers=layers, bn=bn, **kw)


@pytest.mark.parametrize("kind,aggr,algo,feats,dd", [
    ("MPNNConv", "max", "knn", ["relative_position"], "X"),
    ("MPNNConv", "add", "radius", ["point_pair_features"], "X"),
    ("RadarPointGNNConv", "max", "knn", ["relative_position", "relative_velocity"], "XV"),
])
def test_pipeline_matches_oracle(ops, kind, aggr, algo, feats, dd):
    frames = [synthetic.radar_frame(n, seed=10 + s) for s, n in enumerate([300, 250, 1, 180])]
    X, V, ptr = synthetic.frame_batch(frames)
    n = X.shape[0]
    de = go.edge_feature_width(feats)
    c0 = 16
    params = _stack_params(3, c0, 32, de, kind, seed=3)
    x0 = synthetic.node_embeddings(n, c0, seed=1)
    cfg = _pipeline_cfg(ops, params, 3, kind, aggr, algorithm=algo, k=6, r=3.0, distance_definition=dd,
                        edge_features=feats, edge_mode="directed")
    ei, ea, h = ops.pipeline_forward(cfg, _dev(X, torch.float32), _dev(V, torch.float32), _dev(x0), ptr)
    basis = [np.concatenate([f.X_cc, f.V_cc_compensated], axis=1) if dd == "XV" else f.X_cc for f in frames]
    E = go.batched_edges(basis, algo, k=6, r=3.0)
    np.testing.assert_array_equal(ei.cpu().numpy().T, E)                      # bit-exact
    ef = go.edge_features(X, V, E, feats, "directed").astype(np.float32)
    np.testing.assert_allclose(ea.cpu().numpy(), ef, rtol=2e-7, atol=2e-5)
    want = mo.conv_stack_forward(params, torch.from_numpy(x0), torch.from_numpy(E.T.copy()), torch.from_numpy(ef),
                                 3, kind, aggr, dtype=torch.float64)
    assert mo.relative_error(h.cpu(), want) <= 1e-4                           # north-star tolerance
    # host-buffer entry point gives the same bytes
    ei2, ea2, h2 = ops.pipeline_forward_host(cfg, X, V, x0, ptr)
    np.testing.assert_array_equal(ei2, ei.cpu().numpy())
    np.testing.assert_array_equal(ea2, ea.cpu().numpy())
    np.testing.assert_array_equal(h2, h.cpu().numpy())


def test_pipeline_config2_10k(ops):
    """BASELINE config 2: 10 k points, k = 16, 4 x MPNNConv(64 -> 64) + BN + ReLU."""
    fr = synthetic.uniform_square(10_000, seed=0)
    params = _stack_params(4, 64, 64, 2, "MPNNConv", seed=0)
    x0 = synthetic.node_embeddings(10_000, 64)
    cfg = _pipeline_cfg(ops, params, 4, "MPNNConv", "max", algorithm="knn", k=16)
    ei, ea, h = ops.pipeline_forward(cfg, _dev(fr.X_cc, torch.float32), _dev(fr.V_cc_compensated, torch.float32), _dev(x0))
    E = go.knn_edges_sklearn(fr.X_cc, 16)   # the reference's own sklearn call
    assert go.kth_gap_is_tie_free(fr.X_cc, E, 16)
    np.testing.assert_array_equal(ei.cpu().numpy().T, E)
    ef = go.edge_features(fr.X_cc, fr.V_cc_compensated, E, ["relative_position"], "directed").astype(np.float32)
    np.testing.assert_array_equal(ea.cpu().numpy(), ef)
    want = mo.conv_stack_forward(params, torch.from_numpy(x0), torch.from_numpy(E.T.copy()), torch.from_numpy(ef),
                                 4, "MPNNConv", "max", dtype=torch.float64)
    assert mo.relative_error(h.cpu(), want) <= 1e-4


def test_pipeline_headline_100k_against_reference_calls(ops):
    """The headline shape (100 k points, k = 16, 1.6 M edges, 4 x MPNNConv(64 -> 64) + BN + ReLU):
    edge_index bit-exact against the reference's own sklearn call, edge_attr exact, embeddings within
    the north star's 1e-4 against the fp32 CPU restatement of the PyG ops."""
    fr = synthetic.uniform_square(100_000, seed=0)
    params = _stack_params(4, 64, 64, 2, "MPNNConv", seed=0)
    x0 = synthetic.node_embeddings(100_000, 64)
    cfg = _pipeline_cfg(ops, params, 4, "MPNNConv", "max", algorithm="knn", k=16)
    ei, ea, h = ops.pipeline_forward(cfg, _dev(fr.X_cc, torch.float32), _dev(fr.V_cc_compensated, torch.float32), _dev(x0))
    E = go.knn_edges_sklearn(fr.X_cc, 16)
    assert go.kth_gap_is_tie_free(fr.X_cc, E, 16)
    np.testing.assert_array_equal(ei.cpu().numpy().T, E)
    ef = go.edge_features(fr.X_cc, fr.V_cc_compensated, E, ["relative_position"], "directed").astype(np.float32)
    np.testing.assert_array_equal(ea.cpu().numpy(), ef)
    with torch.no_grad():
        want = mo.conv_stack_forward(params, torch.from_numpy(x0), torch.from_numpy(E.T.copy()), torch.from_numpy(ef),
                                     4, "MPNNConv", "max", dtype=torch.float32)
    assert mo.relative_error(h.cpu(), want) <= 1e-4
    # run-to-run determinism: the same bytes again
    ei2, ea2, h2 = ops.pipeline_forward(cfg, _dev(fr.X_cc, torch.float32), _dev(fr.V_cc_compensated, torch.float32), _dev(x0))
    assert torch.equal(ei, ei2) and torch.equal(ea, ea2) and torch.equal(h, h2)


def test_pipeline_config3_shape_batched_frames(ops):
    """BASELINE config 3 shape: 64 RadarScenes-like frames x 300 points, k = 20, d = 128 (two layers here);
    P = 258 exceeds the tensor-core tile, so this exercises the fp32 CUDA-core contraction path."""
    frames = [synthetic.radar_frame(300, seed=s) for s in range(64)]
    X, V, ptr = synthetic.frame_batch(frames)
    n = X.shape[0]
    params = _stack_params(2, 128, 128, 2, "MPNNConv", seed=5)
    x0 = synthetic.node_embeddings(n, 128, seed=2)
    cfg = _pipeline_cfg(ops, params, 2, "MPNNConv", "max", algorithm="knn", k=20)
    ei, ea, h = ops.pipeline_forward(cfg, _dev(X, torch.float32), _dev(V, torch.float32), _dev(x0), ptr)
    E = go.batched_edges([f.X_cc for f in frames], "knn", k=20, backend="sklearn")
    np.testing.assert_array_equal(ei.cpu().numpy().T, E)
    ef = go.edge_features(X, V, E, ["relative_position"], "directed").astype(np.float32)
    want = mo.conv_stack_forward(params, torch.from_numpy(x0), torch.from_numpy(E.T.copy()), torch.from_numpy(ef),
                                 2, "MPNNConv", "max", dtype=torch.float64)
    assert mo.relative_error(h.cpu(), want) <= 1e-4


def test_pipeline_config4_shape_point_pair_features(ops):
    """BASELINE config 4 shape (scaled down): nuScenes-like frames of 2 000 points, k = 20, rotation-invariant
    point-pair features (De = 4), 4 x MPNNConv(64 -> 64)."""
    frames = [synthetic.nuscenes_frame(2000, seed=s) for s in range(4)]
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
    want = mo.conv_stack_forward(params, torch.from_numpy(x0), torch.from_numpy(E.T.copy()), ea.cpu(),
                                 4, "MPNNConv", "max", dtype=torch.float64)
    assert mo.relative_error(h.cpu(), want) <= 1e-4


def test_tensor_core_path_isolated_nodes_and_weight_cache(ops):
    """The tcgen05 contraction folds W_m W_t into the update weights; nodes without incoming edge must
    still get post_mlp([x ; 0]) -- and the cached weight images must follow in-place weight updates."""
    g = torch.Generator().manual_seed(11)
    n, e, c = 900, 6000, 64
    x = torch.randn(n, c, generator=g)
    ei = torch.randint(0, n // 2, (2, e), generator=g)      # half of the nodes never receive a message
    ea = torch.randn(e, 2, generator=g)
    p = 2 * c + 2
    params = {"pre_mlp.0.weight": torch.randn(p, p, generator=g) / p ** 0.5, "pre_mlp.0.bias": torch.randn(p, generator=g) * 0.1,
              "post_mlp.0.weight": torch.randn(c, p + c, generator=g) / (p + c) ** 0.5, "post_mlp.0.bias": torch.randn(c, generator=g) * 0.1}
    dev = {k: v.to(DEV) for k, v in params.items()}
    cp = ops.ConvParams("MPNNConv", c, c, 2, "max", [(dev["pre_mlp.0.weight"], dev["pre_mlp.0.bias"])],
                        [(dev["post_mlp.0.weight"], dev["post_mlp.0.bias"])])
    csc = ops.csc_build(ei.to(DEV), n)
    got = ops.conv_forward(cp, x.to(DEV), csc, ea.to(DEV)).cpu()
    want = mo.mpnn_conv_forward(params, x, ei, ea, "max", dtype=torch.float64)
    assert mo.relative_error(got, want) <= 2e-5
    isolated = torch.arange(n // 2, n)
    exact = torch.nn.functional.linear(torch.cat([x[isolated], torch.zeros(len(isolated), p)], 1).double(),
                                       params["post_mlp.0.weight"].double(), params["post_mlp.0.bias"].double())
    assert mo.relative_error(got[isolated], exact) <= 2e-5
    assert cp._packed is not None                                     # the tensor-core images were cached
    with torch.no_grad():                                             # in-place update bumps the version counter
        dev["post_mlp.0.weight"].mul_(0.5)
    params["post_mlp.0.weight"] = params["post_mlp.0.weight"] * 0.5
    got2 = ops.conv_forward(cp, x.to(DEV), csc, ea.to(DEV)).cpu()
    want2 = mo.mpnn_conv_forward(params, x, ei, ea, "max", dtype=torch.float64)
    assert mo.relative_error(got2, want2) <= 2e-5


def test_pipeline_split_layout_radius_isolated_nodes_and_frames(ops):
    """C = 64 MPNNConv (split message layout, TMA / bulk-copy contractions) on a radius graph over batched
    radar frames: many nodes without incoming edge, in-degrees from 0 to far above 32, a one-point frame."""
    frames = [synthetic.radar_frame(n, seed=40 + s) for s, n in enumerate([280, 1, 330, 150])]
    Xs = [f.X_cc.copy() for f in frames]
    Vs = [f.V_cc_compensated.copy() for f in frames]
    blob = np.random.default_rng(7)                    # a dense blob: in-degrees of ~60 inside frame 0
    Xs[0] = np.concatenate([Xs[0], np.array([40.0, 10.0]) + 0.5 * blob.standard_normal((60, 2))])
    Vs[0] = np.concatenate([Vs[0], blob.standard_normal((60, 2))])
    X, V = np.concatenate(Xs), np.concatenate(Vs)
    ptr = np.zeros(len(Xs) + 1, dtype=np.int64)
    ptr[1:] = np.cumsum([x.shape[0] for x in Xs])
    n = X.shape[0]
    feats = ["relative_position"]
    params = _stack_params(2, 64, 64, 2, "MPNNConv", seed=9)
    x0 = synthetic.node_embeddings(n, 64, seed=2)
    for aggr in ("max", "mean"):
        cfg = _pipeline_cfg(ops, params, 2, "MPNNConv", aggr, algorithm="radius", k=6, r=2.0, distance_definition="X",
                            edge_features=feats, edge_mode="directed")
        ei, ea, h = ops.pipeline_forward(cfg, _dev(X, torch.float32), _dev(V, torch.float32), _dev(x0), ptr)
        E = go.batched_edges(Xs, "radius", k=6, r=2.0)
        np.testing.assert_array_equal(ei.cpu().numpy().T, E)
        indeg = np.bincount(E[:, 1], minlength=n)
        assert (indeg == 0).any() and indeg.max() > 32
        ef = go.edge_features(X, V, E, feats, "directed").astype(np.float32)
        want = mo.conv_stack_forward(params, torch.from_numpy(x0), torch.from_numpy(E.T.copy()), torch.from_numpy(ef),
                                     2, "MPNNConv", aggr, dtype=torch.float64)
        assert mo.relative_error(h.cpu(), want) <= 1e-4


def test_host_entry_point_replays_its_graph(ops):
    """rgnn_pipeline_forward_host captures its work into a CUDA graph and replays it while the arguments stay
    the same: a replay must read the buffers' current contents, and changed frame offsets must re-capture."""
    import ctypes as C
    from radargnn_b200 import _lib
    lib = _lib.load()
    dev = torch.device(DEV)
    fr = synthetic.uniform_square(3000, seed=4)
    params = _stack_params(2, 64, 64, 2, "MPNNConv", seed=5)
    cfg = _pipeline_cfg(ops, params, 2, "MPNNConv", "max", algorithm="knn", k=8, edge_features=["relative_position"])
    handle = ops._PipelineHandle(cfg)
    n = 3000
    pos_h = torch.from_numpy(fr.X_cc.astype(np.float32)).pin_memory()
    vel_h = torch.from_numpy(fr.V_cc_compensated.astype(np.float32)).pin_memory()
    x0_h = torch.from_numpy(synthetic.node_embeddings(n, 64, seed=6)).pin_memory()
    h_h = torch.empty((n, 64), dtype=torch.float32).pin_memory()
    sp = torch.cuda.current_stream().cuda_stream

    one = np.array([0, n], dtype=np.int64)
    two = np.array([0, 1200, n], dtype=np.int64)    # other frame offsets: another graph
    need = max(lib.rgnn_pipeline_host_workspace_bytes(C.byref(handle.desc), n, len(p) - 1, ops.knn_edge_count(p, 8), 64)
               for p in (one, two))
    ws = _lib.workspace(need, dev)                  # kept alive: its pointer is part of the replay key

    def rerun(ptr):
        n_edges = ops.knn_edge_count(ptr, 8)
        _lib.check(lib.rgnn_pipeline_forward_host(C.byref(handle.desc), pos_h.data_ptr(), vel_h.data_ptr(), x0_h.data_ptr(), 64,
                                                  ptr.ctypes.data, len(ptr) - 1, None, n_edges, None, h_h.data_ptr(),
                                                  ws.data_ptr(), ws.numel(), sp))
        return h_h.clone()

    def device_path(ptr):
        _, _, h = ops.pipeline_forward(cfg, pos_h.to(dev), vel_h.to(dev), x0_h.to(dev), ptr)
        return h.cpu()

    a = rerun(one)                                  # first call: capture + launch
    torch.testing.assert_close(a, device_path(one), rtol=0, atol=0)
    b = rerun(one)
    torch.testing.assert_close(b, a, rtol=0, atol=0)
    x0_h.mul_(0.5).add_(0.25)                       # new contents, same pointers: the replay must see them
    c = rerun(one)
    torch.testing.assert_close(c, device_path(one), rtol=0, atol=0)
    assert not torch.equal(c, a)
    d = rerun(two)
    torch.testing.assert_close(d, device_path(two), rtol=0, atol=0)


@pytest.mark.parametrize("depth", [1, 2, 3])
def test_host_pipeline_matches_single_calls(ops, depth):
    """rgnn_pipeline_submit_host / rgnn_pipeline_wait_host with several batches in flight: every batch gives the bytes
    of a synchronous rgnn_pipeline_forward_host call, in submission order, slots and replayed graphs reused, batches
    of changing size (re-capture) and an empty batch included."""
    params = _stack_params(2, 64, 64, 2, "MPNNConv", seed=7)
    cfg = _pipeline_cfg(ops, params, 2, "MPNNConv", "max", algorithm="knn", k=8, edge_features=["relative_position"])
    sizes = [2500, 2500, 2500, 1800, 2500, 0, 700, 2500, 2500]
    batches = []
    for i, n in enumerate(sizes):
        fr = synthetic.uniform_square(max(n, 1), seed=20 + i)
        X, V = fr.X_cc[:n].astype(np.float32), fr.V_cc_compensated[:n].astype(np.float32)
        ptr = np.array([0, n // 3, n], dtype=np.int64) if n > 30 else np.array([0, n], dtype=np.int64)
        batches.append((X, V, synthetic.node_embeddings(max(n, 1), 64, seed=40 + i)[:n], ptr))
    want = [ops.pipeline_forward_host(cfg, *b) for b in batches]
    pipe = ops.HostPipeline(cfg, depth=depth)
    got = []
    for b in batches:
        if pipe.full:
            got.append(pipe.wait(copy=True))
        pipe.submit(*b)
    assert pipe.in_flight == min(depth, len(batches))
    while pipe.in_flight:
        got.append(pipe.wait(copy=True))
    assert len(got) == len(want)
    for (ei, ea, h), (ei0, ea0, h0) in zip(got, want):
        np.testing.assert_array_equal(ei, ei0)
        np.testing.assert_array_equal(ea, ea0)
        np.testing.assert_array_equal(h, h0)


def test_host_pipeline_slot_errors(ops):
    """Slot bookkeeping of the C-ABI: waiting on an idle slot, submitting to a busy one, slots out of range and an
    error raised by the data (non-finite coordinates) reported by the wait."""
    import ctypes as C
    from radargnn_b200 import _lib
    lib = _lib.load()
    assert lib.rgnn_pipeline_wait_host(0) == 1            # RGNN_ERR_INVALID_ARGUMENT: nothing submitted
    assert lib.rgnn_pipeline_wait_host(-1) == 1 and lib.rgnn_pipeline_wait_host(_lib.HOST_SLOTS) == 1
    params = _stack_params(1, 64, 64, 2, "MPNNConv", seed=8)
    cfg = _pipeline_cfg(ops, params, 1, "MPNNConv", "max", algorithm="knn", k=4, edge_features=["relative_position"])
    fr = synthetic.uniform_square(500, seed=3)
    x0 = synthetic.node_embeddings(500, 64, seed=2)
    pipe = ops.HostPipeline(cfg, depth=2)
    pipe.submit(fr.X_cc, fr.V_cc_compensated, x0)
    handle = ops._PipelineHandle(cfg)
    ptr = np.array([0, 500], dtype=np.int64)
    slot = pipe._slots[0]
    busy = lib.rgnn_pipeline_submit_host(0, C.byref(handle.desc), slot["pos"].data_ptr(), slot["vel"].data_ptr(),
                                         slot["x0"].data_ptr(), 64, ptr.ctypes.data, 1, None, 2000, None,
                                         slot["h"].data_ptr(), slot["ws"].data_ptr(), slot["ws"].numel(), None)
    assert busy == 1                                       # slot 0 has not been waited for
    pipe.wait()
    bad = fr.X_cc.copy()
    bad[17, 0] = np.nan
    pipe.submit(bad, fr.V_cc_compensated, x0)
    with pytest.raises(ValueError):                        # sklearn check_array: "Input contains NaN"
        pipe.wait()
    pipe.submit(fr.X_cc, fr.V_cc_compensated, x0)          # the slot is usable again
    _, _, h = pipe.wait()
    assert np.isfinite(h).all()


@pytest.mark.parametrize("aggr", ["max", "mean", "min"])
def test_fused_layer_degenerate_sizes(ops, aggr):
    """Fused aggregate + update kernel (C = 64 MPNNConv) on degenerate graphs: a single node, no edges at all, fewer
    rows than one 4-row unit, one row past a tile boundary, an in-degree of 100 per node."""
    g = torch.Generator().manual_seed(0)
    c, de = 64, 2
    p = 2 * c + de
    params = {"pre_mlp.0.weight": torch.randn(p, p, generator=g) / p ** 0.5, "pre_mlp.0.bias": torch.randn(p, generator=g) * 0.1,
              "post_mlp.0.weight": torch.randn(c, p + c, generator=g) / (p + c) ** 0.5, "post_mlp.0.bias": torch.randn(c, generator=g) * 0.1}
    cp = _conv_params(ops, params, {"kind": "MPNNConv", "aggr": aggr})
    for n, e in ((1, 0), (1, 3), (3, 0), (2, 5), (17, 40), (129, 1000), (4097, 9), (300, 30000)):
        x = torch.randn(n, c, generator=g)
        ei = torch.randint(0, n, (2, e), generator=g)
        ea = torch.randn(e, de, generator=g)
        want = mo.mpnn_conv_forward(params, x, ei, ea, aggr, dtype=torch.float64)
        got = ops.conv_forward(cp, x.to(DEV), ops.csc_build(ei.to(DEV), n), ea.to(DEV)).cpu()
        assert mo.relative_error(got, want) <= 2e-5 and torch.isfinite(got).all(), (n, e)
