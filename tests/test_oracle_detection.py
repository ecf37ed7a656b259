"""CPU: the detection oracle (loss, NMS, nearest neighbour, time index) against what the reference itself calls --
torch's loss modules, torchvision.ops.nms, sklearn -- and against the reference's known answer for rotated NMS."""
import numpy as np
import pytest
import torch

from oracle import detection_oracle as do


def test_loss_oracle_equals_vectorised_torch():
    g = torch.Generator().manual_seed(0)
    n, k, nb = 200, 6, 5
    cls, bb = torch.randn(n, k, generator=g), torch.randn(n, nb, generator=g) * 2
    y = torch.cat([torch.randint(0, k, (n, 1), generator=g).float(), torch.randn(n, nb, generator=g) * 2], dim=1)
    w = torch.rand(k, generator=g) + 0.5
    loss, lc, lb, num = do.detection_loss(cls, bb, y, w, bg_index=5, alpha=0.7, beta=1.3)
    fg = y[:, 0].long() != 5
    want_bb = torch.nn.functional.huber_loss(bb[fg], y[fg, 1:], reduction="none").mean(dim=1).sum() / fg.sum()
    want_cls = torch.nn.functional.cross_entropy(cls, y[:, 0].long(), weight=w)
    assert num == int(fg.sum())
    assert lc == pytest.approx(float(want_cls), rel=1e-6)
    assert lb == pytest.approx(float(want_bb), rel=1e-5)
    assert loss == pytest.approx(0.7 * float(want_cls) + 1.3 * float(want_bb), rel=1e-5)


def test_nms_aligned_oracle_equals_torchvision():
    torchvision = pytest.importorskip("torchvision")
    g = torch.Generator().manual_seed(1)
    for n in (1, 7, 150):
        xy = torch.rand(n, 2, generator=g) * 20 - 5       # some negative coordinates: the reference shifts them
        wh = torch.rand(n, 2, generator=g) * 6 + 0.1
        boxes = torch.cat([xy, xy + wh], dim=1)
        scores = torch.rand(n, generator=g)
        for thr in (0.1, 0.5):
            shift = abs(float(boxes.min())) + 100 if float(boxes.min()) < 0 else 0      # postprocessing.py:400-404
            want = torchvision.ops.nms((boxes + shift).float(), scores.float(), thr).numpy()
            got = do.nms_aligned(boxes.numpy(), scores.numpy(), thr)
            np.testing.assert_array_equal(got, want)


def test_nms_rotated_reference_known_answer():
    """reference test/test_postprocessor.py:8-35"""
    box_matrix = np.array([[1, 2, 1, 1, 90], [1, 2.9, 1, 1, 90]], dtype=np.float64)
    scores = np.array([0.2, 0.7])
    iou = (0.1 * 1) / ((1 + 1) - (0.1 * 1))
    assert do.iou_rotated(box_matrix[0], box_matrix[1]) == pytest.approx(iou, rel=1e-12)
    np.testing.assert_array_equal(do.nms_rotated(box_matrix, scores, iou - 0.01), [1])
    np.testing.assert_array_equal(do.nms_rotated(box_matrix, scores, iou + 0.01), [1, 0])


def test_iou_rotated_against_axis_aligned_and_identity():
    a = np.array([3.0, 4.0, 2.0, 6.0, 0.0])
    assert do.iou_rotated(a, a) == pytest.approx(1.0, rel=1e-12)
    b = np.array([4.0, 4.0, 2.0, 6.0, 180.0])            # shifted by 1 in x, rotated by half a turn: same rectangle
    assert do.iou_rotated(a, b) == pytest.approx((1 * 6) / (24 - 6), rel=1e-12)
    c = np.array([3.0, 4.0, 6.0, 2.0, 90.0])             # the same rectangle as a, described the other way round
    assert do.iou_rotated(a, c) == pytest.approx(1.0, rel=1e-9)
    d = np.array([30.0, 4.0, 2.0, 6.0, 33.0])
    assert do.iou_rotated(a, d) == 0.0


def test_time_index_is_dense_rank():
    ts = np.array([5.0, 1.0, 5.0, 3.0, 1.0, 9.0])
    np.testing.assert_array_equal(do.time_index(ts), [2, 0, 2, 1, 0, 3])


def _golden(name):
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", name))


def test_oracle_matches_committed_golden_fixtures():
    """tests/golden/detection_loss.npz / nms_aligned.npz (oracle/make_golden.py::detection_fixtures: the trainer's loss
    arithmetic as written, torchvision.ops.nms itself)."""
    d = _golden("detection_loss.npz")
    got = do.detection_loss(torch.from_numpy(d["cls"]), torch.from_numpy(d["bb"]), torch.from_numpy(d["y"]),
                            torch.from_numpy(d["weight"]), int(d["bg_index"]), float(d["alpha"]), float(d["beta"]))
    assert got[0] == pytest.approx(float(d["loss"]), rel=1e-6) and got[3] == int(d["num_bb"])
    m = _golden("nms_aligned.npz")
    for t in (10, 30, 60):
        np.testing.assert_array_equal(do.nms_aligned(m["boxes"], m["scores"], t / 100), m[f"keep_{t}"])
