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


# ---- the reference's own import paths (BASELINE.json north star: "drop in under src/gnnradarobjectdetection") ----
def test_reference_import_paths_resolve_to_the_cuda_backed_classes():
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from gnnradarobjectdetection.gnn.mpnn_layers import MPNNConv, RadarPointGNNConv\n"
        "from gnnradarobjectdetection.gnn.gnn_models import DetNetBasic, get_mlp\n"
        "from gnnradarobjectdetection.gnn.configs import GNNArchitectureConfig\n"
        "from gnnradarobjectdetection.graph_constructor.graph import Graph, GeometricGraph\n"
        "from gnnradarobjectdetection.graph_constructor.features import get_En_equivariant_point_pair_metrics\n"
        "from gnnradarobjectdetection.preprocessor.configs import GraphConstructionConfiguration\n"
        "from gnnradarobjectdetection.preprocessor.radar_point_cloud import RadarPointCloud\n"
        "from gnnradarobjectdetection.preprocessor.radarscenes.dataset_creation import GraphConstructor, create_graph_data\n"
        "from gnnradarobjectdetection.preprocessor.nuscenes.conversion import build_geometric_graph\n"
        "from gnnradarobjectdetection.postprocessor.postprocessing import BoxSuppressor\n"
        "import radargnn_b200.gnn.mpnn_layers as impl, radargnn_b200.ops as ops\n"
        "assert MPNNConv is impl.MPNNConv and RadarPointGNNConv is impl.RadarPointGNNConv\n"
        "import inspect; assert 'ops.conv_forward' in inspect.getsource(MPNNConv.forward)\n"
        "assert 'ops.nms' in inspect.getsource(BoxSuppressor.keep_indices)\n"
        "print('ok')\n") % os.path.join(root, "src")
    # run from another directory: the package must find radargnn_b200 by itself
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr


def test_layers_subclass_the_real_pyg_message_passing_when_it_is_importable(tmp_path):
    """The reference's layers subclass torch_geometric.nn.MessagePassing (gnn/mpnn_layers.py:4,11).  PyG is
    not installed here, so a minimal stand-in package proves the wiring: with torch_geometric importable the
    CUDA-backed layers must BE MessagePassing modules, without it they are plain torch modules."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = tmp_path / "torch_geometric"
    (pkg / "nn").mkdir(parents=True)
    (pkg / "__init__.py").write_text("__version__ = 'stub'\n")
    (pkg / "nn" / "__init__.py").write_text(
        "import torch\n"
        "class MessagePassing(torch.nn.Module):\n"
        "    def __init__(self, aggr='add', flow='source_to_target', node_dim=-2):\n"
        "        super().__init__()\n"
        "        self.aggr, self.flow, self.node_dim = aggr, flow, node_dim\n")
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import torch_geometric.nn as tgnn\n"
        "from radargnn_b200.gnn import MPNNConv, RadarPointGNNConv\n"
        "from radargnn_b200.gnn import _message_passing as mp\n"
        "assert mp.USES_PYG_BASE\n"
        "c = MPNNConv(2, 4, 3, aggr='max'); r = RadarPointGNNConv(2, 1)\n"
        "assert isinstance(c, tgnn.MessagePassing) and isinstance(r, tgnn.MessagePassing)\n"
        "assert c.aggr == 'max' and c.conv_params().aggr == 'max' and c.pre_mlp[0].weight.shape == (7, 7)\n"
        "print('ok')\n") % (root, str(tmp_path))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stderr
    from radargnn_b200.gnn import _message_passing as mp
    if not mp.USES_PYG_BASE:   # this interpreter has no PyG: plain torch module
        from radargnn_b200.gnn import MPNNConv
        assert isinstance(MPNNConv(2, 4, 3), torch.nn.Module)


def test_conv_params_are_cached_per_module_and_follow_reassigned_weights():
    from radargnn_b200.gnn import MPNNConv
    conv = MPNNConv(2, 4, 3)
    a = conv.conv_params()
    assert conv.conv_params() is a                      # same tensors installed: same object (keeps packed images)
    conv.post_mlp[0].weight = torch.nn.Parameter(torch.zeros_like(conv.post_mlp[0].weight))
    b = conv.conv_params()
    assert b is not a and float(b.post[0][0].abs().sum()) == 0.0
    conv.invalidate_packed_weights()
    assert conv.conv_params() is not b
    with torch.inference_mode():
        t = torch.ones(3)
    from radargnn_b200 import ops
    assert ops._version_of(t) == -1                     # no version counter under inference_mode: must not raise


def test_graph_edge_list_setter_drops_stale_device_state():
    from radargnn_b200.graph_constructor.graph import Graph
    g = Graph()
    g._edge_index_dev, g._dev_in_E_order, g._A = object(), True, object()
    g.E = np.array([[0, 1], [1, 0]])
    assert g._edge_index_dev is None and g._A is None and not g._built_edges_current()
    g._n = 2
    assert g.A.tolist() == [[0.0, 1.0], [1.0, 0.0]]
