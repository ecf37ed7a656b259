"""GPU parity of the callers either side of the hot path through the C ABI: loss, NMS,
nearest neighbour, time index, collate -- each against the CPU oracle (oracle/detection_oracle.py), which is itself
pinned on torch / torchvision / sklearn / the reference's known answers (tests/test_oracle_detection.py)."""
import numpy as np
import pytest
import torch

from oracle import detection_oracle as do
from radargnn_b200 import synthetic

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


@pytest.fixture(scope="module")
def ops():
    from radargnn_b200 import ops as _ops
    return _ops


@pytest.mark.parametrize("n,k,nb,bg", [(1, 2, 5, 1), (257, 6, 5, 5), (5000, 6, 5, 5), (300, 4, 3, 0)])
def test_detection_loss_matches_trainer_arithmetic(ops, n, k, nb, bg):
    g = torch.Generator().manual_seed(n)
    cls, bb = torch.randn(n, k, generator=g) * 3, torch.randn(n, nb, generator=g) * 2
    y = torch.cat([torch.randint(0, k, (n, 1), generator=g).float(), torch.randn(n, nb, generator=g) * 2], dim=1)
    w = torch.rand(k, generator=g) + 0.5
    want = do.detection_loss(cls, bb, y, w, bg, alpha=0.6, beta=1.4)
    out = ops.detection_loss(cls.to(DEV), bb.to(DEV), y.to(DEV), w.to(DEV), bg, 0.6, 1.4).cpu().numpy()
    np.testing.assert_allclose(out[:3], want[:3], rtol=2e-6, atol=1e-7)   # the reference sums in fp32
    assert out[3] == want[3] and out[4] == 0
    again = ops.detection_loss(cls.to(DEV), bb.to(DEV), y.to(DEV), w.to(DEV), bg, 0.6, 1.4).cpu().numpy()
    assert np.array_equal(out, again)                                      # deterministic


def test_detection_loss_edge_cases(ops):
    g = torch.Generator().manual_seed(3)
    n, k, nb = 64, 3, 5
    cls, bb = torch.randn(n, k, generator=g), torch.randn(n, nb, generator=g)
    y = torch.cat([torch.full((n, 1), 2.0), torch.full((n, nb), float("nan"))], dim=1)   # background only: no box term
    want = do.detection_loss(cls, bb, y, None, 2)
    out = ops.detection_loss(cls.to(DEV), bb.to(DEV), y.to(DEV), None, 2).cpu().numpy()
    np.testing.assert_allclose(out[:3], want[:3], rtol=2e-6)
    assert out[2] == 0.0 and out[3] == 0
    y[0, 0] = 0.0                                                                          # one foreground node with NaN targets
    out = ops.detection_loss(cls.to(DEV), bb.to(DEV), y.to(DEV), None, 2, nan_to_zero=True).cpu().numpy()
    assert out[2] == 0.0 and np.isfinite(out[0])                                           # trainer.py:206-216
    out = ops.detection_loss(cls.to(DEV), bb.to(DEV), y.to(DEV), None, 2, nan_to_zero=False).cpu().numpy()
    assert np.isnan(out[2])
    y[1, 0] = 7.0                                                                          # label outside [0, K)
    out = ops.detection_loss(cls.to(DEV), bb.to(DEV), y.to(DEV), None, 2).cpu().numpy()
    assert out[4] == 1


@pytest.mark.parametrize("n", [1, 2, 65, 700])
def test_nms_aligned_matches_torchvision_semantics(ops, n):
    g = torch.Generator().manual_seed(n)
    xy = torch.rand(n, 2, generator=g) * 30 - 8
    wh = torch.rand(n, 2, generator=g) * 6 + 0.1
    boxes = torch.cat([xy, xy + wh], dim=1)
    scores = torch.rand(n, generator=g)
    scores[n // 2] = scores[0]                      # a score tie: lower index first
    for thr in (0.05, 0.3, 0.7):
        want = do.nms_aligned(boxes.numpy(), scores.numpy(), thr)
        got = ops.nms(boxes.to(DEV), scores.to(DEV), thr).cpu().numpy()
        np.testing.assert_array_equal(got, want)
    try:
        import torchvision
        shift = abs(float(boxes.min())) + 100 if float(boxes.min()) < 0 else 0
        want_tv = torchvision.ops.nms((boxes + shift).float(), scores.float(), 0.3).numpy()
        if len(np.unique(scores.numpy())) == n:
            np.testing.assert_array_equal(ops.nms(boxes.to(DEV), scores.to(DEV), 0.3).cpu().numpy(), want_tv)
    except ImportError:
        pass


def test_nms_rotated_reference_known_answer_and_random(ops):
    from radargnn_b200.postprocessor import BoxSuppressor
    box_matrix = np.array([[1, 2, 1, 1, 90], [1, 2.9, 1, 1, 90]], dtype=np.float64)   # reference test/test_postprocessor.py:8-35
    scores = np.array([[0.2], [0.7]])
    iou = (0.1 * 1) / ((1 + 1) - (0.1 * 1))
    np.testing.assert_array_equal(BoxSuppressor.keep_indices(box_matrix, scores, iou - 0.01, True).cpu().numpy(), [1])
    np.testing.assert_array_equal(BoxSuppressor.keep_indices(box_matrix, scores, iou + 0.01, True).cpu().numpy(), [1, 0])
    m, s, lab = BoxSuppressor.apply_nms(box_matrix, scores, np.array([3, 4]), iou - 0.01, True)
    assert m.shape == (1, 5) and s.shape == (1, 1) and lab.tolist() == [[4]]
    rng = np.random.default_rng(0)
    n = 300
    boxes = np.column_stack([rng.uniform(-20, 20, n), rng.uniform(-20, 20, n), rng.uniform(0.5, 8, n), rng.uniform(0.5, 4, n),
                             rng.uniform(-180, 180, n)])
    sc = rng.uniform(0, 1, n)
    for thr in (0.1, 0.4):
        want = do.nms_rotated(boxes, sc, thr)
        got = ops.nms(torch.from_numpy(boxes).to(DEV), torch.from_numpy(sc).to(DEV), thr, rotated=True).cpu().numpy()
        np.testing.assert_array_equal(got, want)


def test_nms_frames_do_not_suppress_each_other(ops):
    boxes = torch.tensor([[0, 0, 2, 2], [0, 0, 2, 2], [0.1, 0, 2, 2], [5, 5, 6, 6]], dtype=torch.float32)
    scores = torch.tensor([0.9, 0.8, 0.7, 0.6])
    frame = torch.tensor([0, 1, 1, 0], dtype=torch.int32)
    got = ops.nms(boxes.to(DEV), scores.to(DEV), 0.5, box_frame=frame.to(DEV)).cpu().numpy()
    np.testing.assert_array_equal(got, [0, 3, 1])     # frame 0: boxes 0, 3; frame 1: box 1 suppresses box 2
    assert ops.nms(boxes[:0].to(DEV), scores[:0].to(DEV), 0.5).numel() == 0


def test_nearest_neighbor_matches_sklearn(ops):
    from radargnn_b200.postprocessor import nearest_neighbor_positions
    frames = [synthetic.radar_frame(n, seed=s) for s, n in enumerate([300, 150, 2])]
    X, V, ptr = synthetic.frame_batch(frames)
    idx, pts = ops.nearest_neighbor(torch.from_numpy(X).to(DEV), ptr)
    off = 0
    for f in frames:
        want_idx, want_pts = do.nearest_neighbor_positions(f.X_cc)
        m = f.X_cc.shape[0]
        np.testing.assert_array_equal(idx[off:off + m].cpu().numpy(), want_idx + off)
        np.testing.assert_array_equal(pts[off:off + m].cpu().numpy(), want_pts)
        off += m
    np.testing.assert_array_equal(nearest_neighbor_positions(frames[0].X_cc), do.nearest_neighbor_positions(frames[0].X_cc)[1])


def test_time_index_matches_reference_loop(ops):
    rng = np.random.default_rng(1)
    sizes = [0, 1, 300, 2500]
    ts = [rng.choice(np.array([0.0, 0.05, 0.1, 0.15, 17.5, 1e9]) + f, size=m) for f, m in enumerate(sizes)]
    ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    got = ops.time_index(torch.from_numpy(np.concatenate(ts)).to(DEV), ptr).cpu().numpy()
    want = np.concatenate([do.time_index(t) if t.size else t for t in ts])
    np.testing.assert_array_equal(got, want)


def test_time_index_large_frames(ops):
    """Frames above 8192 points (the BASELINE frames of 10 k / 100 k / 125 k points) take the global-memory path:
    few distinct stamps (radar sweeps), all-distinct stamps, sizes that are no power of two, short frames and an
    empty frame in the same batch."""
    rng = np.random.default_rng(3)
    sizes = [10_000, 0, 300, 100_000, 8193, 1, 20_011]
    ts = []
    for f, m in enumerate(sizes):
        if f == 4:
            ts.append(rng.permutation(m).astype(np.float64) * 0.25)                       # every stamp distinct
        elif f == 6:
            ts.append(np.floor(rng.uniform(0, 5000, size=m)))                              # thousands of ties
        else:
            ts.append(rng.choice(np.array([0.0, 0.05, 0.1, 0.15, 17.5, 1.6e15]) + f, size=m))
    ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    got = ops.time_index(torch.from_numpy(np.concatenate(ts)).to(DEV), ptr).cpu().numpy()
    want = np.concatenate([np.unique(t, return_inverse=True)[1].astype(np.float64) if t.size else t for t in ts])
    np.testing.assert_array_equal(got, want)
    small = do.time_index(ts[2])                                                           # the reference loop agrees
    np.testing.assert_array_equal(got[ptr[2]:ptr[3]], small)


def test_collate_offsets_match_disjoint_union(ops):
    g = torch.Generator().manual_seed(2)
    nodes, edges = [5, 0, 7, 3], [9, 0, 20, 2]
    parts = [torch.randint(0, max(n, 1), (2, e), generator=g) for n, e in zip(nodes, edges)]
    node_ptr = np.concatenate([[0], np.cumsum(nodes)])
    edge_ptr = np.concatenate([[0], np.cumsum(edges)])
    want = torch.cat([p + int(o) for p, o in zip(parts, node_ptr[:-1])], dim=1)          # PyG Batch.from_data_list
    ei, batch = ops.collate_offsets(torch.cat(parts, dim=1).contiguous().to(DEV), edge_ptr, node_ptr)
    assert torch.equal(ei.cpu(), want)
    assert batch.cpu().tolist() == [f for f, n in enumerate(nodes) for _ in range(n)]


def test_collate_graph_data_equals_per_graph_forward(ops):
    """Disjoint-union batching on the device: a layer on the collated batch == the layer on every graph."""
    from radargnn_b200.preprocessor import GraphConstructionConfiguration, GraphConstructor, RadarPointCloud, collate_graph_data, create_graph_data
    cfg_g = GraphConstructionConfiguration("knn", {"k": 5, "r": 1.0}, ["rcs", "time_index", "degree"], ["relative_position"], "directed", "X")
    graphs = []
    for s, n in enumerate([120, 40, 77]):
        fr = synthetic.radar_frame(n, seed=s)
        pc = RadarPointCloud()
        pc.X_cc, pc.V_cc_compensated, pc.rcs, pc.timestamp = fr.X_cc, fr.V_cc_compensated, fr.rcs, fr.timestamp
        g = GraphConstructor.build_geometric_graph(cfg_g, pc)
        np.testing.assert_array_equal(g.F["time_index"].reshape(-1), do.time_index(np.asarray(fr.timestamp).reshape(-1)))
        graphs.append(create_graph_data(g).to(DEV))
    batch = collate_graph_data(graphs)
    assert batch.x.shape[0] == 237 and batch.edge_index.shape[1] == 237 * 5
    off = 0
    for g in graphs:
        m = g.x.shape[0]
        sel = (batch.batch == graphs.index(g))
        assert int(sel.sum()) == m
        cols = (batch.edge_index[0] >= off) & (batch.edge_index[0] < off + m)
        assert torch.equal(batch.edge_index[:, cols] - off, g.edge_index)
        off += m


def test_detection_kernels_match_committed_golden_fixtures(ops):
    """The CUDA loss against the trainer's arithmetic, the CUDA NMS against torchvision.ops.nms (tests/golden/*.npz)."""
    import os
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    d = np.load(os.path.join(gdir, "detection_loss.npz"))
    out = ops.detection_loss(torch.from_numpy(d["cls"]).to(DEV), torch.from_numpy(d["bb"]).to(DEV), torch.from_numpy(d["y"]).to(DEV),
                             torch.from_numpy(d["weight"]).to(DEV), int(d["bg_index"]), float(d["alpha"]), float(d["beta"])).cpu().numpy()
    np.testing.assert_allclose(out[:3], [float(d["loss"]), float(d["loss_cls"]), float(d["loss_bb"])], rtol=2e-6)
    assert out[3] == int(d["num_bb"])
    m = np.load(os.path.join(gdir, "nms_aligned.npz"))
    for t in (10, 30, 60):
        got = ops.nms(torch.from_numpy(m["boxes"]).to(DEV), torch.from_numpy(m["scores"]).to(DEV), t / 100).cpu().numpy()
        np.testing.assert_array_equal(got, m[f"keep_{t}"])
