"""Host-side logic that needs no GPU: the mirrored configuration / module structure, parameter
names (state_dict compatibility with the reference), frame sharding, and the 2-rank loss all-reduce
on the gloo backend."""
import os
import socket

import numpy as np
import pytest
import torch

from helpers import DETNET_FIXTURES, load_module_fixture


def test_graph_construction_configuration():
    from radargnn_b200.preprocessor import GraphConstructionConfiguration
    c = GraphConstructionConfiguration("knn", {"k": 20, "r": 3}, ["rcs"], ["relative_position"], "directed", "X")
    assert c.k == 20 and c.r is None
    c = GraphConstructionConfiguration("radius", {"k": 20, "r": 3}, ["rcs"], ["relative_position"], "directed", "X")
    assert c.r == 3 and c.k is None
    with pytest.raises(Exception, match="Invalid graph construction algorithm selected"):
        GraphConstructionConfiguration("ball", {}, [], [], "directed", "X")


def test_layer_structure_matches_reference_tests():
    # reference test/test_gnn.py:42-116 (structure only; numerics are GPU tests)
    from radargnn_b200.gnn import MPNNConv, RadarPointGNNConv, get_mlp
    conv = RadarPointGNNConv(2, 1, "max", 2, 1)
    assert len(conv.pre_mlp) == 3 and len(conv.post_mlp) == 1
    assert conv.pre_mlp[0].weight.shape == (3, 3) and conv.post_mlp[0].weight.shape == (2, 5)
    conv = MPNNConv(2, 4, 3, post_layers=2)
    assert len(conv.pre_mlp) == 1 and len(conv.post_mlp) == 3
    assert conv.pre_mlp[0].weight.shape == (7, 7) and conv.post_mlp[0].weight.shape == (4, 9)
    conv = MPNNConv(1, 4, 2, use_edge_encoder=True)
    assert conv.pre_mlp[0].weight.shape[1] == 3 and conv.edge_encoder.weight.shape == (1, 2)
    mlp = get_mlp(2, 3, [5], False)
    assert mlp[0].weight.shape == (5, 2) and mlp[2].weight.shape == (3, 5)
    mlp = get_mlp(4, 3, [8, 6], True)
    assert [type(m).__name__ for m in mlp] == ["Linear", "BatchNorm", "ReLU", "Linear", "BatchNorm", "ReLU", "Linear"]


@pytest.mark.parametrize("name", DETNET_FIXTURES)
def test_reference_state_dicts_load_by_name(name):
    from radargnn_b200.gnn import DetNetBasic, GNNArchitectureConfig
    params, meta, _ = load_module_fixture(name)
    cfg = GNNArchitectureConfig(
        int(meta["node_feature_dimension"]), int(meta["edge_feature_dimension"]), list(meta["conv_layer_dimensions"]),
        list(meta["classification_head_layer_dimensions"]), list(meta["regression_head_layer_dimensions"]),
        initial_node_feature_embedding=bool(meta.get("initial_node_feature_embedding", False)),
        initial_edge_feature_embedding=bool(meta.get("initial_edge_feature_embedding", False)),
        node_feature_embedding_layer_dimensions=meta.get("node_feature_embedding_layer_dimensions"),
        edge_feature_embedding_layer_dimensions=meta.get("edge_feature_embedding_layer_dimensions"),
        conv_layer_type=meta["conv_layer_type"], batch_norm_in_mlps=bool(meta.get("batch_norm_in_mlps", False)),
        aggregation_function=meta.get("aggregation_function", "max"))
    model = DetNetBasic(cfg)
    assert sorted(model.state_dict().keys()) == sorted(params.keys())
    model.load_state_dict(params, strict=True)
    p = model.convs[0].conv_params()
    assert p.conv_type == meta["conv_layer_type"] and p.aggr == meta.get("aggregation_function", "max")


def test_weights_are_read_at_call_time():
    # the reference's tests re-assign layer.weight after construction (test/test_gnn.py:13-16)
    from radargnn_b200.gnn import MPNNConv
    conv = MPNNConv(2, 4, 3)
    conv.pre_mlp[0].weight = torch.nn.Parameter(torch.ones_like(conv.pre_mlp[0].weight))
    assert float(conv.conv_params().pre[0][0].sum()) == 49.0


def test_shard_frames_balances_points():
    from radargnn_b200.sharding import local_frame_ptr, shard_frames
    sizes = [300] * 64
    shards = shard_frames(sizes, 8)
    assert shards == [(8 * r, 8 * r + 8) for r in range(8)]
    sizes = [100_000, 10, 10, 50_000, 50_000, 1, 99_000]
    shards = shard_frames(sizes, 4)
    assert shards[0][0] == 0 and shards[-1][1] == len(sizes)
    assert all(a[1] == b[0] for a, b in zip(shards, shards[1:]))
    loads = [sum(sizes[a:b]) for a, b in shards]
    assert max(loads) <= 110_020
    assert shard_frames([5, 5], 4)[-1][1] == 2 and sum(b - a for a, b in shard_frames([5, 5], 4)) == 2
    assert shard_frames([], 2) == [(0, 0), (0, 0)]
    ptr = np.array([0, 300, 500, 900, 1000])
    assert local_frame_ptr(ptr, 1, 3).tolist() == [0, 200, 600]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, out):
    import torch.distributed as dist
    from radargnn_b200.sharding import all_reduce_loss, shard_frames
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sizes = [7, 3, 5, 9]
        a, b = shard_frames(sizes, world)[rank]
        values = torch.arange(sum(sizes), dtype=torch.float64)
        lo, hi = sum(sizes[:a]), sum(sizes[:b])
        mean = all_reduce_loss(values[lo:hi].sum(), torch.tensor(float(hi - lo)))
        out[rank] = float(mean)
    finally:
        dist.destroy_process_group()


def test_loss_all_reduce_two_ranks_gloo():
    import torch.multiprocessing as mp
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_rank_main, args=(2, port, out), nprocs=2, join=True)
        total = sum([7, 3, 5, 9])
        assert out[0] == out[1] == pytest.approx((total - 1) / 2)


def test_all_reduce_loss_single_process():
    from radargnn_b200.sharding import all_reduce_loss
    assert float(all_reduce_loss(torch.tensor(6.0), torch.tensor(4.0))) == 1.5
