"""Pin the MPNN oracle: reference known answers (test/test_gnn.py) and golden vectors
produced by the reference's own layer classes (oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import mpnn_oracle as mo
from helpers import CONV_FIXTURES, DETNET_FIXTURES, load_module_fixture


def ones_params(shapes):
    p = {}
    for key, shape in shapes.items():
        p[key + ".weight"] = torch.ones(shape)
        p[key + ".bias"] = torch.zeros(shape[0])
    return p


def test_known_answer_436():
    # reference test/test_gnn.py:119-172 -- MPNNConv(2, 4, 3, post_layers=2, aggr="max"), all weights 1
    p = ones_params({"pre_mlp.0": (7, 7), "post_mlp.0": (4, 9), "post_mlp.2": (4, 4)})
    x = torch.tensor([[1., 1.], [2., 2.]])
    ei = torch.tensor([[0, 1, 0], [1, 0, 1]])
    ea = torch.tensor([[3., 3., 3.], [4., 4., 4.], [1., 1., 1.]])
    out = mo.mpnn_conv_forward(p, x, ei, ea, "max")
    assert out[1].tolist() == [436.0] * 4


def test_known_answer_edge_encoder_23():
    # reference test/test_gnn.py:175-221 -- MPNNConv(1, 4, 2, use_edge_encoder=True)
    p = ones_params({"pre_mlp.0": (3, 3), "post_mlp.0": (4, 4)})
    p["edge_encoder.weight"] = torch.full((1, 2), 2.0)
    p["edge_encoder.bias"] = torch.zeros(1)
    x = torch.tensor([[1.], [2.]])
    ei = torch.tensor([[0, 1], [1, 0]])
    ea = torch.tensor([[1., 1.], [2., 2.]])
    out = mo.mpnn_conv_forward(p, x, ei, ea, "max", use_edge_encoder=True)
    assert out[1, 0].item() == 23.0


def test_known_answer_mlp_shapes():
    # reference test/test_gnn.py:9-25 (get_mlp(2,3,[5]) -> [10,10,10]) and :79-116 (pre -> 7, post -> 72)
    p = ones_params({"m.0": (5, 2), "m.2": (3, 5)})
    assert mo.run_sequential(p, "m", torch.tensor([1., 1.])).tolist() == [10.0, 10.0, 10.0]
    p = ones_params({"pre_mlp.0": (7, 7), "post_mlp.0": (4, 9), "post_mlp.2": (4, 4)})
    assert mo.run_sequential(p, "pre_mlp", torch.ones(1, 7))[0].tolist() == [7.0] * 7
    assert mo.run_sequential(p, "post_mlp", torch.full((1, 9), 2.0))[0].tolist() == [72.0] * 4


def test_zero_in_degree_is_zero_for_every_aggregation():
    m = torch.tensor([[-3.0, 2.0], [-5.0, 1.0]])
    t = torch.tensor([1, 1])
    for aggr, row1 in (("max", [-3.0, 2.0]), ("min", [-5.0, 1.0]), ("add", [-8.0, 3.0]), ("mean", [-4.0, 1.5])):
        out = mo.scatter_aggregate(m, t, 3, aggr)
        assert out[0].tolist() == [0.0, 0.0] and out[2].tolist() == [0.0, 0.0]
        assert out[1].tolist() == row1


@pytest.mark.parametrize("name", CONV_FIXTURES)
def test_conv_matches_reference_golden(name):
    params, meta, d = load_module_fixture(name)
    if meta["kind"] == "MPNNConv":
        out = mo.mpnn_conv_forward(params, d["x"], d["edge_index"], d["edge_attr"], meta["aggr"],
                                   bool(meta.get("use_edge_encoder", False)))
    else:
        out = mo.radar_point_gnn_conv_forward(params, d["x"], d["edge_index"], d["edge_attr"], meta["aggr"])
    assert out.shape == d["out"].shape
    assert mo.relative_error(out, d["out"]) <= 2e-6
    # the fixture exercises zero in-degree nodes (last two nodes receive no message)
    assert int(d["edge_index"][1].max()) < d["x"].shape[0] - 2


@pytest.mark.parametrize("name", DETNET_FIXTURES)
def test_detnet_matches_reference_golden(name):
    params, meta, d = load_module_fixture(name)
    cls, bb = mo.det_net_forward(
        params, d["x"], d["edge_index"], d["edge_attr"], n_layers=len(meta["conv_layer_dimensions"]),
        conv_type=meta["conv_layer_type"], aggr=meta.get("aggregation_function", "max"),
        node_embedding=bool(meta.get("initial_node_feature_embedding", False)),
        edge_embedding=bool(meta.get("initial_edge_feature_embedding", False)))
    assert mo.relative_error(cls, d["out_cls"]) <= 1e-5
    assert mo.relative_error(bb, d["out_bb"]) <= 1e-5


def test_fp64_truth_mode_close_to_fp32():
    params, meta, d = load_module_fixture("mpnn_wide")
    a = mo.mpnn_conv_forward(params, d["x"], d["edge_index"], d["edge_attr"], "max")
    b = mo.mpnn_conv_forward(params, d["x"], d["edge_index"], d["edge_attr"], "max", dtype=torch.float64)
    assert b.dtype == torch.float64 and mo.relative_error(a, b) < 1e-5
