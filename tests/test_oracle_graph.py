"""Pin the graph oracle: reference known answers, golden vectors produced by the
reference's own code (oracle/make_golden.py), sklearn agreement.  CPU only."""
import numpy as np
import pytest

from oracle import graph_oracle as go
from oracle import reference_loop as rl
from oracle import reference_loader
from radargnn_b200 import synthetic
from helpers import GRAPH_FIXTURES, load_graph_fixture, assert_edge_features_close


# ---- known answers of the reference's tests ---------------------------------------------
def test_ppf_known_answer_directed():
    # reference test/test_graph_constructor.py:6-17
    out = go.point_pair_features([[1, 1]], [[3, 2]], [[0, 1]], [[1, 0]], "directed")[0]
    assert [round(v, 2) for v in out] == [2.24, 90.0, 63.43, 26.57]


def test_ppf_known_answer_zero_velocity():
    # reference test/test_graph_constructor.py:20-31
    out = go.point_pair_features([[1, 1]], [[3, 2]], [[0, 1]], [[0, 0]], "directed")[0]
    assert [round(v, 2) for v in out] == [2.24, 90.0, 63.43, 90.0]


def test_edge_row_known_answer():
    # reference test/test_graph_constructor.py:34-59
    X = np.array([[1, 1], [3, 2]]); V = np.array([[0, 1], [1, 0]])
    E = go.knn_edges_bruteforce(X, 1)
    feats = ["point_pair_features", "spatial_euclidean_distance", "velocity_euclidean_distance",
             "relative_position", "relative_velocity"]
    ef = go.edge_features(X, V, E, feats, "directed")
    assert np.round(ef[0], 2).tolist() == [2.24, 90, 63.43, 26.57, 2.24, 1.41, -2, -1, -1, 1]
    # reverse edge: directed PPF is not symmetric
    assert np.round(ef[1], 2).tolist() == [2.24, 90, 153.43, 116.57, 2.24, 1.41, 2, 1, 1, -1]
    loop = rl.edge_feature_loop(X.astype(float), V.astype(float), E, feats, "directed")
    np.testing.assert_allclose(loop, ef, rtol=1e-13, atol=1e-12)


def test_node_row_known_answer():
    # reference test/test_graph_constructor.py:62-88
    X = np.array([[1, 1], [3, 2]]); V = np.array([[0, 1], [1, 0]])
    F = {"rcs": np.array([1.8, 2.6]).reshape(2, 1), "time_index": np.array([100, 101]).reshape(2, 1)}
    E = go.knn_edges_bruteforce(X, 1)
    xf = go.node_features(X, V, F, E, ["rcs", "time_index", "degree", "velocity_vector_length",
                                       "velocity_vector", "spatial_coordinates"])
    assert xf[1].tolist() == [2.6, 101, 1, 1, 1, 0, 3, 2]


def test_graph_constructor_known_answers():
    # reference test/test_preprocessor.py:207-230 and :233-257
    X = np.array([[1, 1], [3, 2], [5, 8]], dtype=float)
    E, ef, xf = go.build_geometric_graph(
        X_cc=X, V_cc=np.ones_like(X), rcs=None, timestamp=np.array([100, 101, 102]).reshape(3, 1),
        algorithm="knn", k=1, r=None, node_feature_names=["spatial_coordinates", "time_index"],
        edge_feature_names=["spatial_euclidean_distance"], edge_mode="directed", distance_definition="X")
    assert E.tolist() == [[0, 1], [1, 0], [2, 1]]
    assert ef[0, 0] == 5 ** 0.5
    assert xf[1].tolist() == [3, 2, 1]
    X = np.array([[1, 1], [2, 2], [10, 10]], dtype=float)
    V = np.ones_like(X); V[0, :] = 100
    kw = dict(X_cc=X, V_cc=V, rcs=None, timestamp=None, algorithm="knn", k=1, r=None,
              node_feature_names=["spatial_coordinates"], edge_feature_names=["spatial_euclidean_distance"],
              edge_mode="directed")
    assert go.build_geometric_graph(distance_definition="X", **kw)[0].tolist() == [[0, 1], [1, 0], [2, 1]]
    assert go.build_geometric_graph(distance_definition="XV", **kw)[0].tolist() == [[0, 1], [1, 2], [2, 1]]


def test_degree_idempotent_and_undirected():
    # reference test/test_graph_constructor.py:91-103 + SURVEY appendix A (degree = |N_out U N_in|)
    E = go.knn_edges_bruteforce(np.array([[1, 1], [3, 2]]), 1)
    assert go.undirected_degree(E, 2).tolist() == [1, 1]
    X = synthetic.radar_frame(60, seed=2).X_cc
    E = go.knn_edges_bruteforce(X, 3)
    A = np.zeros((60, 60)); A[E[:, 0], E[:, 1]] = 1
    np.testing.assert_array_equal(go.undirected_degree(E, 60), ((A + A.T) > 0).sum(1))
    np.testing.assert_array_equal(rl.degree_like_reference(A)[:, 0], go.undirected_degree(E, 60))


def test_k_not_smaller_than_n_raises():
    with pytest.raises(ValueError):
        go.knn_edges_bruteforce(np.random.rand(4, 2), 4)
    assert go.knn_edges_bruteforce(np.random.rand(1, 2), 3).shape == (0, 2)   # graph.py:45


def test_invalid_feature_raises():
    X = np.random.rand(4, 2)
    with pytest.raises(Exception, match="Invalid feature specified"):
        go.edge_features(X, X, go.knn_edges_bruteforce(X, 1), ["nope"], "directed")


def test_dot_product_error_raises():
    with pytest.raises(Exception, match="Error in dot product calculation"):
        go._clamped_dot(np.array([[1.1, 0.0]]), np.array([[1.0, 0.0]]))
    assert go._clamped_dot(np.array([[1.0005, 0.0]]), np.array([[-1.0, 0.0]]))[0] == -1.0


# ---- golden vectors from the reference's own code ---------------------------------------
@pytest.mark.parametrize("name", GRAPH_FIXTURES)
def test_oracle_matches_reference_golden(name):
    fx = load_graph_fixture(name)
    E, ef, xf = go.build_geometric_graph(
        X_cc=fx["X_cc"], V_cc=fx["V_cc"], rcs=fx["rcs"], timestamp=fx["timestamp"],
        algorithm=fx["algorithm"], k=fx["k"], r=fx["r"], node_feature_names=fx["node_features"],
        edge_feature_names=fx["edge_features"], edge_mode=fx["edge_mode"],
        distance_definition=fx["distance_definition"])
    ref_E, ref_ef = fx["E"], fx["E_feat"]
    if fx["algorithm"] == "radius":   # row-internal order is a KD-tree artefact: canonicalise
        ref_E, ref_ef = go.canonicalise_rows(ref_E, ref_ef)
    np.testing.assert_array_equal(E, ref_E)                      # bit-exact incl. k-NN order
    # fp64 agreement (the reference takes 2-norms through an SVD: last-ulp differences only)
    assert_edge_features_close(ef, ref_ef, fx["edge_features"])
    np.testing.assert_allclose(xf, fx["X_feat"], rtol=1e-13, atol=0)
    # the per-edge port used as the timed CPU baseline agrees as well
    with np.errstate(invalid="ignore"):
        loop = rl.edge_feature_loop(fx["X_cc"], fx["V_cc"], ref_E, fx["edge_features"], fx["edge_mode"])
    assert_edge_features_close(loop, ref_ef, fx["edge_features"])


def test_ppf_pairs_golden(golden_dir):
    z = np.load(golden_dir + "/ppf_pairs.npz")
    for mode in ("directed", "undirected"):
        out = go.point_pair_features(z["p1"], z["p2"], z["v1"], z["v2"], mode)
        np.testing.assert_allclose(out, z[mode], rtol=1e-12, atol=1e-6, equal_nan=True)
        np.testing.assert_array_equal(np.isnan(out), np.isnan(z[mode]))
    # degenerate rows really are in the fixture: zero velocity -> 90 deg, coincident -> 90 deg
    assert np.all(z["directed"][0:4, 1] == 90.0) and np.all(z["directed"][12:16, 2] == 90.0)


# ---- the restatement against sklearn (the reference's third-party search) ----------------
@pytest.mark.parametrize("n,k,dims,seed", [(300, 16, 2, 0), (300, 20, 2, 1), (257, 5, 4, 2), (12, 3, 2, 3),
                                           (9, 4, 2, 4), (1000, 16, 2, 5)])
def test_knn_bruteforce_equals_sklearn(n, k, dims, seed):
    rng = np.random.default_rng(seed)
    X = rng.uniform(0, np.sqrt(n), (n, dims)).astype(np.float32).astype(np.float64)
    a, b = go.knn_edges_bruteforce(X, k), go.knn_edges_sklearn(X, k)
    assert go.kth_gap_is_tie_free(X, a, k)
    np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("n,r,dims,seed", [(300, 3.0, 2, 0), (300, 1.5, 2, 1), (200, 4.0, 4, 2), (9, 2.0, 2, 3)])
def test_radius_bruteforce_equals_sklearn(n, r, dims, seed):
    rng = np.random.default_rng(seed)
    X = rng.uniform(0, np.sqrt(n), (n, dims)).astype(np.float32).astype(np.float64)
    a = go.radius_edges_bruteforce(X, r)
    b = go.canonicalise_rows(go.radius_edges_sklearn(X, r))
    np.testing.assert_array_equal(a, b)


def test_batched_edges_offsets():
    frames = [synthetic.radar_frame(40, seed=s).X_cc for s in range(3)] + [np.zeros((1, 2))]
    E = go.batched_edges(frames, "knn", k=3)
    assert E.shape == (120 * 3, 2)
    assert E[:120].max() < 40 and E[120:240].min() >= 40 and E[240:].min() >= 80


@pytest.mark.skipif(not reference_loader.available(), reason="needs /root/reference (builder container only)")
def test_vectorised_oracle_equals_reference_loop_random():
    gr = reference_loader.load("graph_constructor.graph")
    for seed in range(3):
        fr = synthetic.radar_frame(30, seed=40 + seed, extent=(20.0, 20.0))
        for mode in ("directed", "undirected"):
            g = gr.GeometricGraph(); g.X, g.V = fr.X_cc, fr.V_cc_compensated
            g.build(fr.X_cc, "knn", k=4)
            feats = list(go.EDGE_FEATURE_WIDTH)
            with np.errstate(invalid="ignore"):
                g.extract_node_pair_features(feats, mode)
            np.testing.assert_array_equal(g.E, go.knn_edges_bruteforce(fr.X_cc, 4))
            ours = go.edge_features(fr.X_cc, fr.V_cc_compensated, g.E, feats, mode)
            assert_edge_features_close(ours, g.E_feat, feats)
