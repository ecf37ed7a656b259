"""Shared helpers for the parity tests (loading golden fixtures, metrics)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

GRAPH_FIXTURES = [
    "graph_knn3_directed_all", "graph_knn5_undirected_all", "graph_knn4_xv",
    "graph_radius3_config1", "graph_knn20_shipped", "graph_knn16_uniform",
    "graph_knn4_tiny_brute", "graph_radius4_tiny_brute",
]
CONV_FIXTURES = ["mpnn_max", "mpnn_add", "mpnn_mean", "mpnn_deep", "mpnn_encoder", "mpnn_wide",
                 "rpgnn_max", "rpgnn_add_deep"]
DETNET_FIXTURES = ["detnet_mpnn", "detnet_rpgnn"]


def load_graph_fixture(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files}
    d["algorithm"] = str(d["algorithm"])
    d["edge_mode"] = str(d["edge_mode"])
    d["distance_definition"] = str(d["distance_definition"])
    d["node_features"] = [str(s) for s in d["node_features"]]
    d["edge_features"] = [str(s) for s in d["edge_features"]]
    d["k"] = None if int(d["k"]) < 0 else int(d["k"])
    d["r"] = None if float(d["r"]) < 0 else float(d["r"])
    return d


def load_module_fixture(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    params = {k[len("param::"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param::")}
    meta = {}
    for k in z.files:
        if k.startswith("meta::"):
            v = z[k]
            meta[k[len("meta::"):]] = v.tolist() if v.ndim else v.item()
    data = {k: torch.from_numpy(z[k]) for k in z.files if "::" not in k}
    return params, meta, data


def f32_ulp_distance(a, b):
    """Distance in float32 units-in-the-last-place between two float32 arrays."""
    a = np.ascontiguousarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    a = np.where(a < 0, np.int64(-2**31) - a, a)
    b = np.where(b < 0, np.int64(-2**31) - b, b)
    return np.abs(a - b)


def angle_column_mask(features):
    """True for the columns of E_feat that hold arccos-derived angles (degrees)."""
    mask = []
    for f in features:
        if f == "point_pair_features":
            mask += [False, True, True, True]
        elif f in ("relative_position", "relative_velocity"):
            mask += [False, False]
        else:
            mask += [False]
    return np.array(mask, dtype=bool)


def assert_edge_features_close(actual, expected, features, rtol=1e-12, angle_atol=2e-6):
    """fp64 comparison of edge-feature matrices.  Distances / differences agree to
    rounding (the reference takes 2-norms through an SVD: last-ulp differences);
    angles go through arccos, which is ill-conditioned near 0 and 180 degrees
    (d(theta) = d(cos)/sin(theta)), hence the absolute tolerance in degrees."""
    actual, expected = np.asarray(actual, dtype=np.float64), np.asarray(expected, dtype=np.float64)
    assert actual.shape == expected.shape
    np.testing.assert_array_equal(np.isnan(actual), np.isnan(expected))
    ang = angle_column_mask(features)
    np.testing.assert_allclose(actual[:, ~ang], expected[:, ~ang], rtol=rtol, atol=1e-12, equal_nan=True)
    np.testing.assert_allclose(actual[:, ang], expected[:, ang], rtol=rtol, atol=angle_atol, equal_nan=True)
