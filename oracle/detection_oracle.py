"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the callers either side of the hot path.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

* ``detection_loss``: the arithmetic of reference gnn/trainer.py:184-231 (train) / :276-298 (validation), literally:
  torch CrossEntropyLoss(weight) + the per-node Python loop over torch HuberLoss.
* ``nms_aligned`` / ``nms_rotated``: greedy suppression in descending score order
  (postprocessor/postprocessing.py:336-435).  The aligned variant is pinned against torchvision.ops.nms (the
  function the reference calls, :408; installed here) in tests/test_oracle_detection.py; the rotated one restates
  detectron2's nms_rotated (absent here: not vendored by the reference, no wheel) and is pinned on the reference's
  own known answer (test/test_postprocessor.py:8-35) -- beyond that, parity unpinned.
* ``nearest_neighbor_positions``: dataset_creation.py:314-318 with sklearn's kneighbors_graph itself.
* ``time_index``: dataset_creation.py:214-223, the loop as written."""
from __future__ import annotations

import numpy as np
import torch


def detection_loss(cls, bb, y, class_weight, bg_index, alpha=1.0, beta=1.0, delta=1.0, nan_to_zero=True):
    """Returns (loss, loss_cls, loss_bb, num_bb) as Python floats; tensors are CPU torch tensors."""
    cross_entropy = torch.nn.CrossEntropyLoss(weight=class_weight)
    huber = torch.nn.HuberLoss(delta=delta)
    label_true = y[:, 0].long()
    bb_true = y[:, 1:]
    loss_cls = cross_entropy(cls, label_true)
    loss_bb = 0
    num_bb = 0
    for i, label in enumerate(label_true):          # trainer.py:190-199
        if label != bg_index:
            num_bb += 1
            loss_bb = loss_bb + huber(bb_true[i, :], bb[i, :])
    loss_bb = loss_bb / num_bb if num_bb != 0 else 0
    if nan_to_zero and isinstance(loss_bb, torch.Tensor) and np.isnan(loss_bb.item()):
        loss_bb = 0
    loss = alpha * loss_cls + beta * loss_bb
    return float(loss), float(loss_cls), float(loss_bb), num_bb


def _shift(m, cols):
    mn = np.min(m[:, cols])
    return (abs(mn) + 100) if mn < 0 else 0


def nms_aligned(boxes, scores, thr, shift_negative=True):
    """boxes [n, 4] (x1, y1, x2, y2); fp32 arithmetic like torchvision's kernel.  Kept indices, descending score."""
    b = np.asarray(boxes, dtype=np.float32).copy()
    s = np.asarray(scores, dtype=np.float32).reshape(-1)
    if shift_negative and b.shape[0]:
        b = (b + np.float32(_shift(b, slice(0, 4)))).astype(np.float32)
    order = sorted(range(len(s)), key=lambda i: (-s[i], i))
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    keep, dead = [], np.zeros(len(s), dtype=bool)
    for pos, i in enumerate(order):
        if dead[i]:
            continue
        keep.append(i)
        for j in order[pos + 1:]:
            if dead[j]:
                continue
            w = max(np.float32(min(b[i, 2], b[j, 2]) - max(b[i, 0], b[j, 0])), np.float32(0))
            h = max(np.float32(min(b[i, 3], b[j, 3]) - max(b[i, 1], b[j, 1])), np.float32(0))
            inter = np.float32(w * h)
            if np.float32(inter / np.float32(np.float32(area[i] + area[j]) - inter)) > np.float32(thr):
                dead[j] = True
    return np.asarray(keep, dtype=np.int64)


def _corners(box):
    cx, cy, w, h, a = box
    t = np.deg2rad(a)
    c2, s2 = np.cos(t) * 0.5, np.sin(t) * 0.5
    p0 = (cx + s2 * h + c2 * w, cy + c2 * h - s2 * w)
    p1 = (cx - s2 * h + c2 * w, cy - c2 * h - s2 * w)
    return [p0, p1, (2 * cx - p0[0], 2 * cy - p0[1]), (2 * cx - p1[0], 2 * cy - p1[1])]


def iou_rotated(a, b):
    """IoU of two (cx, cy, w, h, degrees) boxes: convex polygon clipping + shoelace area, fp64."""
    area_a, area_b = a[2] * a[3], b[2] * b[3]
    if area_a < 1e-14 or area_b < 1e-14:
        return 0.0
    sx, sy = (a[0] + b[0]) * 0.5, (a[1] + b[1]) * 0.5
    pa = _corners((a[0] - sx, a[1] - sy, a[2], a[3], a[4]))
    pb = _corners((b[0] - sx, b[1] - sy, b[2], b[3], b[4]))
    cross = lambda u, v: u[0] * v[1] - u[1] * v[0]
    orient = 1.0 if cross((pb[1][0] - pb[0][0], pb[1][1] - pb[0][1]), (pb[2][0] - pb[1][0], pb[2][1] - pb[1][1])) >= 0 else -1.0
    poly = list(pa)
    for e in range(4):
        c0, c1 = pb[e], pb[(e + 1) % 4]
        edge = (c1[0] - c0[0], c1[1] - c0[1])
        out = []
        for i in range(len(poly)):
            p, q = poly[i], poly[(i + 1) % len(poly)]
            dp = orient * cross(edge, (p[0] - c0[0], p[1] - c0[1]))
            dq = orient * cross(edge, (q[0] - c0[0], q[1] - c0[1]))
            if dp >= 0:
                out.append(p)
            if (dp >= 0) != (dq >= 0):
                t = dp / (dp - dq)
                out.append((p[0] + t * (q[0] - p[0]), p[1] + t * (q[1] - p[1])))
        poly = out
        if not poly:
            return 0.0
    if len(poly) < 3:
        return 0.0
    area2 = sum(cross(poly[i], poly[(i + 1) % len(poly)]) for i in range(len(poly)))
    inter = abs(area2) * 0.5
    return inter / (area_a + area_b - inter)


def nms_rotated(boxes, scores, thr, shift_negative=True):
    b = np.asarray(boxes, dtype=np.float64).copy()
    s = np.asarray(scores, dtype=np.float64).reshape(-1)
    if shift_negative and b.shape[0]:
        b[:, :2] += _shift(b, slice(0, 2))
    order = sorted(range(len(s)), key=lambda i: (-s[i], i))
    keep, dead = [], np.zeros(len(s), dtype=bool)
    for pos, i in enumerate(order):
        if dead[i]:
            continue
        keep.append(i)
        for j in order[pos + 1:]:
            if not dead[j] and iou_rotated(b[i], b[j]) > thr:
                dead[j] = True
    return np.asarray(keep, dtype=np.int64)


def nearest_neighbor_positions(X):
    from sklearn.neighbors import kneighbors_graph
    A_full = kneighbors_graph(X, 1, mode='connectivity', include_self=False).toarray()
    idx = np.where(A_full == 1)[1]
    return idx, X[idx]


def time_index(timestamp):
    timestamps = np.unique(timestamp)
    t_idx = np.zeros_like(timestamp)
    for i, _ in enumerate(timestamps):
        idx = np.where(timestamp == timestamps[i])[0]
        t_idx[idx] = int(i)
    return t_idx
